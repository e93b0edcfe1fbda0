"""The sharded generation loop on real GPUs: N ranks (NCCL) must reproduce the 1-rank run exactly,
because Philox draws are indexed by GLOBAL offspring number and selection is replicated."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def scheme(rank, world):
    from genomicsimulation_b200 import SimData, sharding, synthetic
    ds = synthetic.founders(64, 6000, 7, 120.0, seed=4)
    sim = SimData.from_dataset(ds, device=rank)
    be = sharding.EngineBackend(sim)
    lo, hi = sharding.offspring_range(64, rank, world)
    pop = sharding.ShardedPopulation(be, 64, np.arange(lo, hi, dtype=np.uint32))
    out = []
    for gen in range(3):
        chosen, pool = pop.select_parents(eff=0, top_n=12)
        pop = pop.next_generation(pool, 500 + 7 * gen, seed=77, generation=gen)
        out.append((chosen.tolist(), pop.gather_bvs(0).cpu().numpy().tolist()))
    sim.close()
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        q.put((rank, scheme(rank, world)))
    finally:
        dist.destroy_process_group()


def test_single_rank_scheme_runs():
    out = scheme(0, 1)
    assert len(out) == 3 and len(out[2][1]) == 514
    bv = np.array(out[2][1])
    assert np.isfinite(bv).all() and bv.std() > 0
    # selection moves the mean upward generation over generation
    assert np.mean(out[2][1]) > np.mean(out[0][1])


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    single = scheme(0, 1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=500) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == single and got[1] == single
