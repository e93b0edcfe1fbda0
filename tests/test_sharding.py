"""Host-side plumbing of the sharded generation, on CPU: the index ranges (the C function itself —
it needs no device), and the one thing torch.distributed does for the engine: carrying every rank's
128-byte exchange handle to every rank, here over gloo with world_size 2.  The data path (P2P
pushes, selection, crossing) is covered on the GPU by tests/test_gpu_sharding.py."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from genomicsimulation_b200 import _lib, sharding   # noqa: E402


def test_ranges_partition_the_generation_and_match_the_c_abi():
    E = _lib.engine()
    for n, w in ((10, 3), (7, 8), (1000000, 8), (0, 2), (5, 1), (64, 3), (514, 3)):
        spans = [sharding.offspring_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
        for r in range(w):
            lo, hi = C.c_uint64(), C.c_uint64()
            E.gsb200_shard_range(n, r, w, C.byref(lo), C.byref(hi))
            assert (lo.value, hi.value) == spans[r]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = bytes([rank + 1]) * _lib.SHARD_HANDLE_BYTES
        q.put((rank, sharding.exchange_handles(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_handles_reach_every_rank_in_rank_order():
    assert sharding.exchange_handles(b"\x07" * _lib.SHARD_HANDLE_BYTES) == b"\x07" * _lib.SHARD_HANDLE_BYTES   # no process group: world 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = b"\x01" * _lib.SHARD_HANDLE_BYTES + b"\x02" * _lib.SHARD_HANDLE_BYTES
    assert got[0] == want and got[1] == want
