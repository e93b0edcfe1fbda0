#!/usr/bin/env python
"""BASELINE.json configs[2] as a breeding scheme: G generations of recurrent selection, N_TOTAL offspring per generation
x 50k SNPs, top 1 % by GEBV as parents of the next, sharded over the ranks torchrun starts — through the host mirror
(include/gsb200_gsc.h: gsb_shard_select_and_cross, pedigrees tracked, ids allocated; gsb_shard_get_local_bvs;
gsb_delete_group).  Prints per-generation wall time (max over ranks) and the genetic gain.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/bench_scheme.py [n_total] [generations]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
from genomicsimulation_b200 import SimData, sharding, synthetic   # noqa: E402

N_TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
GENS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F, M, TOP = 1000, 50000, max(2, N_TOTAL // 100)
sim = SimData.from_dataset(synthetic.founders(F, M, 21, 150.0, seed=2), device=local)
sim.use_native_draws(seed=2026)
lo, hi = sharding.offspring_range(N_TOTAL, rank, world)
n = hi - lo
sim.shard_init(rank, world, N_TOTAL, TOP)
sim.shard_connect(sharding.exchange_handles(sim.shard_handle()))
g = sim.make_random_crosses(sim.founders, n, 0, 1)          # generation 0: this rank's slice of N_TOTAL random crosses among the founders
bv = np.empty(n)
times, means = [], []
for gen in range(GENS):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    new = sim.shard_select_and_cross(g, N_TOTAL, 1, TOP, N_TOTAL)
    sim.shard_local_bvs(bv)
    sim.delete_group(g)
    g = new
    sim.shard_finish()
    tt = torch.tensor([time.perf_counter() - t, float(bv.sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        tt[0] = mx[0]
    times.append(float(tt[0])); means.append(float(tt[1]) / N_TOTAL)
if rank == 0:
    print("recurrent selection: %d generations x %d offspring x %d SNPs, top %d selected, %d GPU(s), through the host mirror" % (GENS, N_TOTAL, M, TOP, world))
    for gen, (t, m) in enumerate(zip(times, means)):
        print("  generation %2d: %7.3f ms   mean GEBV of the generation evaluated %.4f" % (gen, t * 1e3, m))
    print("  median %.3f ms per generation = %.3e offspring/s; genetic gain over %d generations %.4f" %
          (float(np.median(times[1:])) * 1e3, N_TOTAL / float(np.median(times[1:])), GENS, means[-1] - means[0]))
sim.close()
if world > 1:
    dist.destroy_process_group()
