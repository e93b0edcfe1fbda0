"""World-size-2 gloo test of the multi-GPU host logic (genomicsimulation_b200/sharding.py) on CPU:
the data plane is a numpy stand-in, the collectives and index bookkeeping are the real ones.
The sharded result must equal the single-rank result (independence from the number of ranks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from genomicsimulation_b200 import sharding   # noqa: E402


class NumpyBackend:
    """Rows are plain byte vectors; a "cross" is a deterministic function of (parents, seed, stream,
    GLOBAL unit) like the engine's Philox stream; a "GEBV" is a weighted byte sum."""

    def __init__(self, width=24):
        self.width, self.rows = width, {}
        self.next = 0
        self.w = np.linspace(-1, 1, width)

    def _new(self, data):
        ids = []
        for d in data:
            self.rows[self.next] = np.asarray(d, dtype=np.uint8)
            ids.append(self.next)
            self.next += 1
        return np.array(ids, dtype=np.uint32)

    def gebv(self, rows, eff):
        return torch.tensor([float(self.rows[int(r)] @ self.w) for r in rows], dtype=torch.float64)

    def gather_rows(self, rows):
        if len(rows) == 0:
            return torch.zeros((0, self.width), dtype=torch.uint8)
        return torch.from_numpy(np.stack([self.rows[int(r)] for r in rows]))

    def adopt_rows(self, packed):
        return self._new(packed.numpy())

    def cross(self, p1, p2, map_index, seed, stream, first_unit):
        out = []
        for i, (a, b) in enumerate(zip(p1, p2)):
            rng = np.random.default_rng([seed, stream, first_unit + i])
            mask = rng.integers(0, 2, self.width).astype(bool)
            out.append(np.where(mask, self.rows[int(a)], self.rows[int(b)]))
        return self._new(out)


def scheme(rank, world):
    be = NumpyBackend()
    n0 = 37
    founders = np.random.default_rng(5).integers(0, 255, (n0, be.width)).astype(np.uint8)
    lo, hi = sharding.offspring_range(n0, rank, world)
    pop = sharding.ShardedPopulation(be, n0, be._new(founders[lo:hi]))
    history = []
    for gen in range(3):
        chosen, pool_rows = pop.select_parents(eff=0, top_n=9)
        pop = pop.next_generation(pool_rows, 41 + gen, seed=11, generation=gen)
        history.append((chosen.tolist(), pop.gather_bvs(0).numpy().round(9).tolist()))
    return history


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, scheme(rank, world)))
    finally:
        dist.destroy_process_group()


def test_offspring_range_and_selection():
    for n, w in ((10, 3), (7, 8), (1000000, 8)):
        spans = [sharding.offspring_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    bv = np.array([1.0, 3.0, 3.0, -2.0, 0.5])
    assert sharding.select_global(bv, 2).tolist() == [1, 2]           # ties by lower index
    assert sharding.select_global(bv, 2, low_is_best=True).tolist() == [3, 4]
    assert sharding.select_global(torch.from_numpy(bv), 2).tolist() == [1, 2]
    # pairs do not depend on which slice of the generation a rank draws
    a, b = sharding.draw_pairs(50, 200000, 9)
    a2, b2 = sharding.draw_pairs(50, 200000, 9, 70000, 140001)
    assert (a[70000:140001] == a2).all() and (b[70000:140001] == b2).all()
    a, b = sharding.draw_pairs(5, 1000, 3)
    assert (a != b).all() and a.min() >= 0 and a.max() <= 4


@pytest.mark.timeout(180)
def test_two_ranks_equal_one_rank():
    single = scheme(0, 1)
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
    assert got[0] == single and got[1] == single
