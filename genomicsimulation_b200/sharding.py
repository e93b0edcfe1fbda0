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
    gsb200_top_n, evaluated identically on every rank."""
    bv_all = np.asarray(bv_all, dtype=np.float64)
    key = bv_all if low_is_best else -bv_all
    order = np.argsort(key, kind="stable")
    return order[:top_n]


def draw_pairs(n_pool, n_offspring, seed):
    """Random parent pairs (a != b) for a whole generation from one host stream, with the
    reference's pick rule round(u * (n - 1)) redrawn on collision (sim-operations.c:8556-8581)."""
    rng = np.random.default_rng(seed)
    a = np.rint(rng.random(n_offspring) * (n_pool - 1)).astype(np.int64)
    b = np.rint(rng.random(n_offspring) * (n_pool - 1)).astype(np.int64)
    clash = a == b
    while clash.any():
        b[clash] = np.rint(rng.random(int(clash.sum())) * (n_pool - 1)).astype(np.int64)
        clash = a == b
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
        bvs = self.gather_bvs(eff).cpu().numpy()
        chosen = select_global(bvs, top_n, low_is_best)
        mine = [(k, int(g)) for k, g in enumerate(chosen) if self.lo <= g < self.hi]
        counts = []
        for r in range(self.world):
            l, h = offspring_range(self.n_total, r, self.world)
            counts.append(int(((chosen >= l) & (chosen < h)).sum()))
        send_rows = np.array([self.local_rows[g - self.lo] for _, g in mine], dtype=np.uint32)
        packed = self.backend.gather_rows(send_rows)                          # tensor [len(mine), row_bytes] u8
        pool = self._all_gather_ragged(packed, counts)                        # rank-major order
        # rank-major position -> selection order
        order = []
        for r in range(self.world):
            l, h = offspring_range(self.n_total, r, self.world)
            order += [k for k, g in enumerate(chosen) if l <= g < h]
        inv = np.empty(len(order), dtype=np.int64)
        inv[np.array(order, dtype=np.int64)] = np.arange(len(order))
        pool = pool[torch.as_tensor(inv, device=pool.device)]
        pool_rows = self.backend.adopt_rows(pool)                             # rows holding the pool here
        return chosen, pool_rows

    def next_generation(self, pool_rows, n_offspring, seed, generation, map_index=0):
        """Every rank makes its own index range of the next generation from the replicated pool."""
        a, b = draw_pairs(len(pool_rows), n_offspring, seed * 1000003 + generation)
        lo, hi = offspring_range(n_offspring, self.rank, self.world)
        rows = self.backend.cross(pool_rows[a[lo:hi]], pool_rows[b[lo:hi]], map_index,
                                  seed=seed, stream=generation, first_unit=lo)
        return ShardedPopulation(self.backend, n_offspring, rows, self.group)


class EngineBackend:
    """CUDA data plane on top of the C-ABI (include/gsb200.h)."""

    def __init__(self, sim):
        self.sim, self.E, self.ctx = sim, sim.E, sim.ctx
        self.row_bytes = self.E.gsb200_store_row_bytes(self.ctx)
        self.next_row = int(self.E.gsb200_store_capacity(self.ctx))
        self.free = []

    def _check(self, st):
        if st != 0:
            raise RuntimeError(self.E.gsb200_last_error().decode())

    def _take(self, n):
        rows = [self.free.pop() for _ in range(min(n, len(self.free)))]
        need = n - len(rows)
        if need:
            self._check(self.E.gsb200_store_reserve(self.ctx, self.next_row + need))
            rows += list(range(self.next_row, self.next_row + need))
            self.next_row += need
        return np.array(rows, dtype=np.uint32)

    def release(self, rows):
        self.free += [int(r) for r in rows]

    def gebv(self, rows, eff):
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.zeros(len(rows), dtype=np.float64)
        self._check(self.E.gsb200_gebv(self.ctx, len(rows), rows.ctypes.data, eff, out.ctypes.data))
        return torch.from_numpy(out).cuda()

    def gather_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        t = torch.empty((len(rows), self.row_bytes), dtype=torch.uint8, device="cuda")
        if len(rows):
            d_rows = torch.from_numpy(rows.astype(np.int32)).cuda()
            torch.cuda.current_stream().synchronize()
            self._check(self.E.gsb200_store_gather_rows_dev(self.ctx, len(rows), d_rows.data_ptr(), t.data_ptr()))
            self._check(self.E.gsb200_synchronize(self.ctx))
        return t

    def adopt_rows(self, packed):
        packed = packed.contiguous()
        torch.cuda.current_stream().synchronize()
        rows = self._take(packed.shape[0])
        if len(rows):
            d_rows = torch.from_numpy(rows.astype(np.int32)).cuda()
            torch.cuda.current_stream().synchronize()
            self._check(self.E.gsb200_store_scatter_rows_dev(self.ctx, len(rows), d_rows.data_ptr(), packed.data_ptr()))
            self._check(self.E.gsb200_synchronize(self.ctx))
        return rows

    def cross(self, p1_rows, p2_rows, map_index, seed, stream, first_unit):
        from ._lib import Draws
        p1 = np.ascontiguousarray(p1_rows, dtype=np.uint32)
        p2 = np.ascontiguousarray(p2_rows, dtype=np.uint32)
        out = self._take(len(p1))
        dr = Draws(mode=0, seed=seed, stream=stream, first_unit=first_unit)
        self._check(self.E.gsb200_cross(self.ctx, len(out), p1.ctypes.data, p2.ctypes.data, map_index, map_index,
                                        out.ctypes.data, C.byref(dr)))
        return out
