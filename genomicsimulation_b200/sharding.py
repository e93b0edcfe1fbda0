"""Multi-GPU generation loop: one process per GPU, offspring sharded by contiguous index range.

Per generation: every rank evaluates its own shard; GEBVs (8 bytes per individual)
are all-gathered; selection is replicated, so every rank agrees on the selected GLOBAL indexes
(ties broken by lower index); the selected individuals' packed rows are all-gathered so every rank
holds the whole parent pool; parent pairs for the WHOLE generation are drawn from one host stream
that is identical on every rank, and each rank makes the offspring of its own index range with
Philox draws indexed by global offspring number.  Results therefore do not depend on the number
of ranks.  Nothing else moves: offspring stay on the GPU that made them.

The data plane is behind a small backend protocol so that the host logic can be exercised on CPU
with the gloo backend (tests/test_sharding.py); `EngineBackend` is the CUDA one.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def offspring_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of rank; the first n_total % world ranks get one more."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def select_global(bv_all, top_n, low_is_best=False):
    """Positions of the top_n best values, best first, ties by lower position — the same rule as
    gsb200_top_n, evaluated identically on every rank.  Accepts a numpy array or a torch tensor
    (a stable sort on whichever device the GEBVs were gathered to)."""
    if isinstance(bv_all, torch.Tensor):
        key = bv_all if low_is_best else -bv_all
        return torch.sort(key, stable=True).indices[:top_n].cpu().numpy()
    bv_all = np.asarray(bv_all, dtype=np.float64)
    key = bv_all if low_is_best else -bv_all
    return np.argsort(key, kind="stable")[:top_n]


_PAIR_BLOCK = 1 << 16


def draw_pairs(n_pool, n_offspring, seed, lo=0, hi=None):
    """Random parent pairs (a != b) for offspring [lo, hi) of a generation of n_offspring, with the
    reference's pick rule round(u * (n - 1)) redrawn on collision (sim-operations.c:8556-8581).
    The host stream is seeded per block of 65 536 offspring, so a rank draws only the blocks that
    overlap its own range and the pairs do not depend on how the generation is sharded."""
    hi = n_offspring if hi is None else hi
    a = np.empty(hi - lo, dtype=np.int64)
    b = np.empty(hi - lo, dtype=np.int64)
    for blk in range(lo // _PAIR_BLOCK, (max(hi, lo + 1) - 1) // _PAIR_BLOCK + 1):
        b_lo, b_hi = blk * _PAIR_BLOCK, min((blk + 1) * _PAIR_BLOCK, n_offspring)
        if b_hi <= b_lo:
            continue
        rng = np.random.default_rng([int(seed) & 0xFFFFFFFFFFFF, blk])
        aa = np.rint(rng.random(b_hi - b_lo) * (n_pool - 1)).astype(np.int64)
        bb = np.rint(rng.random(b_hi - b_lo) * (n_pool - 1)).astype(np.int64)
        clash = aa == bb
        while clash.any():
            bb[clash] = np.rint(rng.random(int(clash.sum())) * (n_pool - 1)).astype(np.int64)
            clash = aa == bb
        s_lo, s_hi = max(lo, b_lo), min(hi, b_hi)
        a[s_lo - lo:s_hi - lo] = aa[s_lo - b_lo:s_hi - b_lo]
        b[s_lo - lo:s_hi - lo] = bb[s_lo - b_lo:s_hi - b_lo]
    return a, b


class ShardedPopulation:
    """The current generation, sharded: this rank holds `local_rows` (backend row handles) of the
    individuals with global indexes [lo, hi)."""

    def __init__(self, backend, n_total, local_rows, group=None):
        self.backend = backend
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.n_total = n_total
        self.lo, self.hi = offspring_range(n_total, self.rank, self.world)
        assert len(local_rows) == self.hi - self.lo
        self.local_rows = np.asarray(local_rows, dtype=np.uint32)

    def _all_gather_ragged(self, t, counts):
        """all_gather of per-rank tensors whose leading sizes are `counts` (known to every rank)."""
        if self.world == 1:
            return t
        width = int(np.prod(t.shape[1:])) if t.dim() > 1 else 1
        cap = max(counts) if counts else 0
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = torch.empty((self.world * cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = [out[r * cap: r * cap + counts[r]] for r in range(self.world)]
        return torch.cat(parts, dim=0)

    def gather_bvs(self, eff):
        """GEBVs of the whole generation, in global index order, on every rank."""
        local = self.backend.gebv(self.local_rows, eff)                       # tensor [hi - lo], f64
        counts = [offspring_range(self.n_total, r, self.world) for r in range(self.world)]
        return self._all_gather_ragged(local, [h - l for l, h in counts])

    def select_parents(self, eff, top_n, low_is_best=False):
        """Replicated selection + parent-pool exchange. Returns (global indexes selected, rows of the
        pool in THIS rank's store, in selection order)."""
        chosen = select_global(self.gather_bvs(eff), top_n, low_is_best)
        bounds = np.array([offspring_range(self.n_total, r, self.world)[0] for r in range(self.world)] + [self.n_total])
        owner = np.searchsorted(bounds, chosen, side="right") - 1
        counts = np.bincount(owner, minlength=self.world).tolist()
        send_rows = self.local_rows[chosen[owner == self.rank] - self.lo]
        packed = self.backend.gather_rows(send_rows)                          # tensor [mine, row_bytes] u8
        pool = self._all_gather_ragged(packed, counts)                        # rank-major order
        # rank-major position p holds selection number perm[p]; put the pool in selection order
        perm = np.argsort(owner, kind="stable")
        inv = np.empty(len(perm), dtype=np.int64)
        inv[perm] = np.arange(len(perm))
        pool = pool[torch.as_tensor(inv, device=pool.device)]
        pool_rows = self.backend.adopt_rows(pool)                             # rows holding the pool here
        return chosen, pool_rows

    def next_generation(self, pool_rows, n_offspring, seed, generation, map_index=0):
        """Every rank makes its own index range of the next generation from the replicated pool."""
        lo, hi = offspring_range(n_offspring, self.rank, self.world)
        pool_rows = np.asarray(pool_rows)
        if hasattr(self.backend, "cross_random"):
            # pairs drawn on the device, indexed by the global offspring number: nothing but the pool's row
            # list goes up, and the generation does not depend on how it is sharded
            rows = self.backend.cross_random(pool_rows, hi - lo, map_index, seed=seed, stream=generation, first_unit=lo)
        else:
            a, b = draw_pairs(len(pool_rows), n_offspring, seed * 1000003 + generation, lo, hi)
            rows = self.backend.cross(pool_rows[a], pool_rows[b], map_index, seed=seed, stream=generation, first_unit=lo)
        return ShardedPopulation(self.backend, n_offspring, rows, self.group)


class EngineBackend:
    """CUDA data plane on top of the C-ABI (include/gsb200.h)."""

    def __init__(self, sim):
        self.sim, self.E, self.ctx = sim, sim.E, sim.ctx
        self.row_bytes = self.E.gsb200_store_row_bytes(self.ctx)
        self.next_row = int(self.E.gsb200_store_capacity(self.ctx))
        self.free = np.empty(0, dtype=np.uint32)

    def _check(self, st):
        if st != 0:
            raise RuntimeError(self.E.gsb200_last_error().decode())

    def _take(self, n):
        k = min(n, len(self.free))
        rows = self.free[len(self.free) - k:]
        self.free = self.free[:len(self.free) - k]
        need = n - k
        if need:
            self._check(self.E.gsb200_store_reserve(self.ctx, self.next_row + need))
            rows = np.concatenate([rows, np.arange(self.next_row, self.next_row + need, dtype=np.uint32)])
            self.next_row += need
        return np.ascontiguousarray(rows, dtype=np.uint32)

    def release(self, rows):
        self.free = np.concatenate([self.free, np.asarray(rows, dtype=np.uint32)])

    def _dev_rows(self, rows):
        return torch.from_numpy(np.ascontiguousarray(rows, dtype=np.uint32).view(np.int32)).cuda()

    def gebv(self, rows, eff):
        """GEBVs of `rows`, left on the device (they go straight into the all-gather)."""
        d_rows = self._dev_rows(rows)
        out = torch.empty(len(rows), dtype=torch.float64, device="cuda")
        torch.cuda.current_stream().synchronize()
        if len(rows):
            self._check(self.E.gsb200_gebv_dev(self.ctx, len(rows), d_rows.data_ptr(), eff, out.data_ptr()))
            self._check(self.E.gsb200_synchronize(self.ctx))
        return out

    def gather_rows(self, rows):
        t = torch.empty((len(rows), self.row_bytes), dtype=torch.uint8, device="cuda")
        if len(rows):
            d_rows = self._dev_rows(rows)
            torch.cuda.current_stream().synchronize()
            self._check(self.E.gsb200_store_gather_rows_dev(self.ctx, len(rows), d_rows.data_ptr(), t.data_ptr()))
            self._check(self.E.gsb200_synchronize(self.ctx))
        return t

    def adopt_rows(self, packed):
        packed = packed.contiguous()
        rows = self._take(packed.shape[0])
        if len(rows):
            d_rows = self._dev_rows(rows)
            torch.cuda.current_stream().synchronize()
            self._check(self.E.gsb200_store_scatter_rows_dev(self.ctx, len(rows), d_rows.data_ptr(), packed.data_ptr()))
            self._check(self.E.gsb200_synchronize(self.ctx))
        return rows

    def cross_random(self, pool_rows, n, map_index, seed, stream, first_unit):
        """n random crosses among `pool_rows` (distinct parents, the reference's pick rule), pairs drawn on
        the device from the stream position of global offspring `first_unit` onwards."""
        from ._lib import Draws
        out = self._take(n)
        if n == 0:
            return out
        pool = np.ascontiguousarray(pool_rows, dtype=np.uint32)
        contiguous = n == 1 or bool(np.all(np.diff(out.astype(np.int64)) == 1))
        dr = Draws(mode=0, seed=seed, stream=stream, first_unit=first_unit)
        self._check(self.E.gsb200_cross_random_pairs(self.ctx, n, 1, pool.ctypes.data, len(pool), None, 0, map_index, map_index,
                                                     None if contiguous else out.ctypes.data, int(out[0]), C.byref(dr), None, None))
        return out

    def cross(self, p1_rows, p2_rows, map_index, seed, stream, first_unit):
        from ._lib import Draws
        n = len(p1_rows)
        out = self._take(n)
        d_p1, d_p2, d_out = self._dev_rows(p1_rows), self._dev_rows(p2_rows), self._dev_rows(out)
        torch.cuda.current_stream().synchronize()
        dr = Draws(mode=0, seed=seed, stream=stream, first_unit=first_unit)
        self._check(self.E.gsb200_cross_dev(self.ctx, n, d_p1.data_ptr(), d_p2.data_ptr(), map_index, map_index,
                                            d_out.data_ptr(), C.byref(dr)))
        self._check(self.E.gsb200_cross_dev_status(self.ctx))      # synchronises
        return out
