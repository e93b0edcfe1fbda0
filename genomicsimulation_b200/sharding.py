"""Sharded recurrent selection — a thin ctypes driver of the C entry points.

The generation loop itself lives behind the C-ABI (include/gsb200.h `gsb200_shard_*`,
csrc/shard.cu) and its host-side mirror (include/gsb200_gsc.h `gsb_shard_*`): GEBV -> push to
every peer -> replicated radix selection -> push of the selected rows into every peer's pool ->
device pair draw -> splice, all on one stream without host synchronisation.  What is left here is
set-up plumbing: `torch.distributed` (any backend) carries the 128-byte exchange handles from
every rank to every rank, once.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SHARD_HANDLE_BYTES


def offspring_range(n_total, rank, world):
    """Contiguous slice [lo, hi) of rank — gsb200_shard_range: the first n_total % world ranks get one more."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def exchange_handles(handle, group=None):
    """All ranks' handles in rank order (bytes of world x SHARD_HANDLE_BYTES), over torch.distributed."""
    import torch
    import torch.distributed as dist
    assert len(handle) == SHARD_HANDLE_BYTES
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bytes(handle)
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
    everyone = torch.empty(world * SHARD_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(everyone, mine, group=group) if dev == "cuda" else \
        dist.all_gather(list(everyone.view(world, SHARD_HANDLE_BYTES).unbind(0)), mine, group=group)
    return everyone.cpu().numpy().tobytes()


class Shard:
    """This rank's shard, on the engine level (gsb200_shard_*): row arrays are the caller's."""

    def __init__(self, sim, rank, world, max_total, max_pool):
        self.sim, self.E, self.ctx = sim, sim.E, sim.ctx
        self.rank, self.world = rank, world
        h = C.c_void_p()
        self._check(self.E.gsb200_shard_create(self.ctx, rank, world, max_total, max_pool, C.byref(h)))
        self.h = h

    def _check(self, st):
        if st != 0:
            raise _lib.EngineError(self.E.gsb200_last_error().decode())

    def handle(self):
        buf = C.create_string_buffer(SHARD_HANDLE_BYTES)
        self._check(self.E.gsb200_shard_handle(self.h, buf))
        return buf.raw

    def connect(self, handles):
        assert len(handles) == self.world * SHARD_HANDLE_BYTES
        self._check(self.E.gsb200_shard_connect(self.h, handles))

    def connect_over(self, group=None):
        self.connect(exchange_handles(self.handle(), group))

    def select(self, local_rows, n_total, eff_index, top_n, low_is_best=False, local_ids=None):
        rows = np.ascontiguousarray(local_rows, dtype=np.uint32)
        ids = None if local_ids is None else np.ascontiguousarray(local_ids, dtype=np.uint32)
        self._keep = (rows, ids)
        self._check(self.E.gsb200_shard_select(self.h, len(rows), rows.ctypes.data, None if ids is None else ids.ctypes.data,
                                               n_total, eff_index, top_n, int(low_is_best)))

    def cross(self, out_rows, map_index, seed, stream, first_unit, picks=False):
        from ._lib import Draws
        out = np.ascontiguousarray(out_rows, dtype=np.uint32)
        dr = Draws(mode=0, seed=seed, stream=stream, first_unit=first_unit)
        k1 = np.empty(len(out), dtype=np.uint32) if picks else None
        k2 = np.empty(len(out), dtype=np.uint32) if picks else None
        self._check(self.E.gsb200_shard_cross(self.h, len(out), map_index, out.ctypes.data, C.byref(dr),
                                              k1.ctypes.data if picks else None, k2.ctypes.data if picks else None))
        return (k1, k2) if picks else None

    def selected(self, ids=False, bvs=False):
        n = self.E.gsb200_shard_pool_size(self.h)
        sel = np.empty(n, dtype=np.uint32)
        sid = np.empty(n, dtype=np.uint32) if ids else None
        bv = np.empty(self.E.gsb200_shard_generation_size(self.h), dtype=np.float64) if bvs else None
        self._check(self.E.gsb200_shard_selected(self.h, sel.ctypes.data, sid.ctypes.data if ids else None, bv.ctypes.data if bvs else None))
        return sel, sid, bv

    def status(self):
        self._check(self.E.gsb200_shard_status(self.h))

    def close(self):
        if self.h:
            self.E.gsb200_shard_destroy(self.h)
            self.h = None
