#!/usr/bin/env python
"""BASELINE.json configs[2] in miniature: G generations of recurrent selection, N_PER_GPU offspring per
generation and GPU x 50k SNPs, top 1 % by GEBV as parents, sharded over the ranks torchrun starts
(genomicsimulation_b200/sharding.py).  Prints per-generation wall time (max over ranks) and the
genetic gain.   torchrun --nproc-per-node N profiles/bench_scheme.py [n_per_gpu] [generations]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomicsimulation_b200 import SimData, sharding, synthetic     # noqa: E402


def main():
    n_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M, F = 50000, 1000
    ds = synthetic.founders(F, M, 21, 150.0, seed=2)
    sim = SimData.from_dataset(ds, device=local)
    be = sharding.EngineBackend(sim)
    n_total = n_per_gpu * world
    # generation 0: every rank makes its share of random crosses from the (replicated) founders
    lo, hi = sharding.offspring_range(n_total, rank, world)
    a, b = sharding.draw_pairs(F, n_total, 12345, lo, hi)
    rows = be.cross(a.astype(np.uint32), b.astype(np.uint32), 0, seed=7, stream=1000, first_unit=lo)
    pop = sharding.ShardedPopulation(be, n_total, rows)
    top = n_total // 100
    times, means = [], []
    for g in range(gens):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        chosen, pool_rows = pop.select_parents(eff=0, top_n=top)
        new = pop.next_generation(pool_rows, n_total, seed=7, generation=g)
        be.release(pop.local_rows)
        be.release(pool_rows)
        pop = new
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
        bv = pop.gather_bvs(0).cpu().numpy()
        means.append(float(bv.mean()))
    if rank == 0:
        per = np.array(times[1:]) if gens > 1 else np.array(times)
        print("recurrent selection: %d GPUs, %d offspring/generation x %d SNPs, top %d selected, pool %.0f MB" %
              (world, n_total, M, top, top * be.row_bytes / 1e6))
        print("seconds per generation (GEBV + all-gather + selection + parent-pool exchange + meiosis): "
              "mean %.4f  min %.4f  -> %.3e offspring/s" % (per.mean(), per.min(), n_total / per.mean()))
        print("mean GEBV by generation: " + " ".join("%.1f" % m for m in means))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
