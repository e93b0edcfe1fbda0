"""configs[3] kernels alone, for ncu: make.doubled.haploids and one selfing generation at N lines x M SNPs.
usage: config4_probe.py [N] [M] [what: dh|self|both] [L2 fetch granularity]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from genomicsimulation_b200 import SimData, synthetic          # noqa: E402
from genomicsimulation_b200._lib import Draws                  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
what = sys.argv[3] if len(sys.argv) > 3 else "both"
F = 100
if len(sys.argv) > 4:          # CU_LIMIT_MAX_L2_FETCH_GRANULARITY (32 / 64 / 128 bytes): how much more than the sectors asked for L2 fetches from DRAM
    torch.zeros(1, device="cuda")          # the primary context, current on this thread
    drv = C.CDLL("libcuda.so.1")
    got = C.c_size_t(0)
    st = drv.cuCtxSetLimit(C.c_int(5), C.c_size_t(int(sys.argv[4])))
    drv.cuCtxGetLimit(C.byref(got), C.c_int(5))
    print("cuCtxSetLimit(MAX_L2_FETCH_GRANULARITY, %s) -> %d, now %d" % (sys.argv[4], st, got.value))
sim = SimData.from_dataset(synthetic.founders(F, M, 21, 150.0, seed=4))
E, ctx = sim.E, sim.ctx


def chk(st):
    if st != 0:
        raise RuntimeError(E.gsb200_last_error().decode())


chk(E.gsb200_store_reserve(ctx, F + 2 * N))
ext = torch.cuda.ExternalStream(E.gsb200_stream(ctx))
founders = np.arange(F, dtype=np.uint32)
chk(E.gsb200_cross_random_pairs(ctx, N, 1, founders.ctypes.data, F, None, 0, 0, 0, None, F, C.byref(Draws(mode=0, seed=4, stream=0, first_unit=0)), None, None))
a = np.arange(F, F + N, dtype=np.uint32)
b = np.arange(F + N, F + 2 * N, dtype=np.uint32)


def timed(fn, reps=3):
    fn(); ext.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext); fn(); e1.record(ext); ext.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


if what in ("self", "both"):
    t = timed(lambda: chk(E.gsb200_self_n_times(ctx, N, a.ctypes.data, 1, 0, b.ctypes.data, C.byref(Draws(mode=0, seed=4, stream=1, first_unit=0)))))
    print("one selfing generation, %d x %d: %.3f ms, %.0f GB/s on M/2 per line" % (N, M, t, N * M / 2 / t / 1e6))
if what in ("dh", "both"):
    t = timed(lambda: chk(E.gsb200_doubled_haploid(ctx, N, a.ctypes.data, 0, b.ctypes.data, C.byref(Draws(mode=0, seed=4, stream=2, first_unit=0)))))
    print("doubled haploids, %d x %d: %.3f ms, %.0f GB/s on 3M/8 per line" % (N, M, t, N * M * 3 / 8 / t / 1e6))
sim.close()
