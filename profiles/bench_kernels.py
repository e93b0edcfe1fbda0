#!/usr/bin/env python
"""Throughput of the hot path's other entry points at the BASELINE.json shapes, through the C-ABI
with host row arrays (copies included).  Prints a table; run on a B200:  python profiles/bench_kernels.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomicsimulation_b200 import SimData, synthetic          # noqa: E402
from genomicsimulation_b200._lib import Draws                   # noqa: E402

PEAK = 6548.8


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    M, CHR, F, N = 50000, 21, 1000, 100000
    ds = synthetic.founders(F, M, CHR, 150.0, seed=2, n_effect_sets=10)
    sim = SimData.from_dataset(ds)
    E, ctx = sim.E, sim.ctx
    assert E.gsb200_store_reserve(ctx, F + 2 * N) == 0
    rng = np.random.default_rng(3)
    p1 = rng.integers(0, F, N).astype(np.uint32)
    p2 = ((p1 + rng.integers(1, F, N)) % F).astype(np.uint32)
    out = np.arange(F, F + N, dtype=np.uint32)
    out2 = np.arange(F + N, F + 2 * N, dtype=np.uint32)
    ptr = lambda a: a.ctypes.data
    dr = Draws(mode=0, seed=1, stream=0, first_unit=0)
    rows = []
    chk = lambda st: (_ for _ in ()).throw(RuntimeError(E.gsb200_last_error().decode())) if st else None

    def line(name, sec, units, alg_bytes, note=""):
        rows.append("%-44s %9.3f ms %12.3e %s/s %8.1f GB/s alg (%.2f of %.0f) %s" %
                    (name, sec * 1e3, units[0] / sec, units[1], alg_bytes / sec / 1e9, alg_bytes / sec / 1e9 / PEAK, PEAK, note))

    t = timeit(lambda: chk(E.gsb200_cross(ctx, N, ptr(p1), ptr(p2), 0, 0, ptr(out), C.byref(dr))))
    line("gsb200_cross 100k x 50k (host rows)", t, (N, "offspring"), N * M / 2)
    t = timeit(lambda: chk(E.gsb200_doubled_haploid(ctx, N, ptr(out), 0, ptr(out2), C.byref(dr))))
    line("gsb200_doubled_haploid 100k x 50k", t, (N, "lines"), N * M / 2)
    t = timeit(lambda: chk(E.gsb200_self_n_times(ctx, N, ptr(out), 5, 0, ptr(out2), C.byref(dr))), reps=3)
    line("gsb200_self_n_times n=5 100k x 50k", t, (N, "lines"), N * M / 2, "(5 generations through HBM scratch)")
    t = timeit(lambda: chk(E.gsb200_clone(ctx, N, ptr(out), ptr(out2))))
    line("gsb200_clone 100k x 50k", t, (N, "lines"), N * M / 2)
    bv = np.zeros(N)
    t = timeit(lambda: chk(E.gsb200_gebv(ctx, N, ptr(out), 0, ptr(bv))))
    line("gsb200_gebv 100k x 50k (host in/out)", t, (N * M, "marker.ind"), N * (M / 4 + 8))
    blk = sim.blocks_evenlength(10, 1)
    off, mk = blk.export()
    n_loc = 20000
    loc = np.zeros((2 * n_loc, len(off) - 1))
    t = timeit(lambda: chk(E.gsb200_local_gebv(ctx, n_loc, ptr(out), len(off) - 1, ptr(off), ptr(mk), 0, ptr(loc))), reps=3)
    line("gsb200_local_gebv 20k x 50k, 210 blocks", t, (n_loc * M, "marker.ind"), n_loc * (M / 4 + 2 * 210 * 8))
    outs = [np.zeros((2 * n_loc, len(off) - 1)) for _ in range(10)]
    ptrs = (C.c_void_p * 10)(*[o.ctypes.data for o in outs])
    effs = np.arange(10, dtype=np.uint32)
    t = timeit(lambda: chk(E.gsb200_local_gebv_multi(ctx, n_loc, ptr(out), len(off) - 1, ptr(off), ptr(mk), 10, ptr(effs), ptrs)), reps=3)
    line("gsb200_local_gebv_multi 20k x 50k, 210 blocks, 10 sets", t, (n_loc * M * 10, "marker.ind.set"), n_loc * (M / 4 + 10 * 2 * 210 * 8),
         "(D2H of 10 x 67 MB included)")
    import torch
    d_outs = [torch.zeros((2 * n_loc, len(off) - 1), dtype=torch.float64, device="cuda") for _ in range(10)]
    dptrs = (C.c_void_p * 10)(*[o.data_ptr() for o in d_outs])
    torch.cuda.synchronize()
    for sets in (1, 4, 10):
        t = timeit(lambda: chk(E.gsb200_local_gebv_multi_dev(ctx, n_loc, ptr(out), len(off) - 1, ptr(off), ptr(mk), sets, ptr(effs), dptrs)), reps=3)
        line("gsb200_local_gebv_multi_dev 20k x 50k, 210 blk, %2d sets" % sets, t, (n_loc * M * sets, "marker.ind.set"),
             n_loc * (M / 4 + sets * 2 * 210 * 8), "(device output; schedule + digit fragments cached on the device)")
    n_cnt = 4000
    cnt = np.zeros((n_cnt, M), dtype=np.uint8)
    t = timeit(lambda: chk(E.gsb200_allele_counts(ctx, n_cnt, ptr(out), b"T", 2, ptr(cnt))), reps=3)
    line("gsb200_allele_counts u8 4k x 50k", t, (n_cnt * M, "cells"), n_cnt * (M / 4 + M), "(output-bound, D2H included)")
    n_io = 2000
    chars = np.zeros((n_io, 2 * M), dtype=np.uint8)
    t = timeit(lambda: chk(E.gsb200_store_download_chars(ctx, n_io, ptr(out), ptr(chars))), reps=3)
    line("gsb200_store_download_chars 2k x 50k", t, (n_io, "genotypes"), n_io * (M / 4 + 2 * M), "(PCIe-bound)")
    t = timeit(lambda: chk(E.gsb200_store_upload_chars(ctx, n_io, ptr(out2), ptr(chars))), reps=3)
    line("gsb200_store_upload_chars 2k x 50k", t, (n_io, "genotypes"), n_io * (M / 4 + 2 * M), "(pageable host buffer: staging copy + PCIe; dictionary built on the device; pinned: profiles/r2_chars_boundary.txt)")
    vals = rng.standard_normal(1000000)
    pos = np.zeros(10000, dtype=np.uint32)
    t = timeit(lambda: chk(E.gsb200_top_n(ctx, len(vals), ptr(vals), len(pos), 0, ptr(pos))))
    line("gsb200_top_n 1M -> top 1% (host values)", t, (len(vals), "values"), len(vals) * 8, "(8 MB of values go up first; the selection itself is 0.07 ms per 1 M, bench.py config3_1gpu select_ms)")
    print("\n".join(rows))
    sim.close()


if __name__ == "__main__":
    main()
