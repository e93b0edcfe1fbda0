#!/usr/bin/env python
"""cross + GEBV as two passes (gsb200_cross_dev, gsb200_gebv_dev) against the fused entry point
(gsb200_cross_gebv_dev), CUDA events on the engine's stream:  python profiles/scripts/fused_ab.py [N] [M] [F]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from genomicsimulation_b200 import SimData, synthetic          # noqa: E402
from genomicsimulation_b200._lib import Draws                   # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    F = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    ds = synthetic.founders(F, M, 21, 150.0, seed=2, n_effect_sets=1)
    sim = SimData.from_dataset(ds)
    E, ctx = sim.E, sim.ctx
    assert E.gsb200_store_reserve(ctx, F + N) == 0
    rng = np.random.default_rng(3)
    p1 = rng.integers(0, F, N)
    p2 = (p1 + rng.integers(1, F, N)) % F
    d_p1, d_p2 = torch.from_numpy(p1.astype(np.int32)).cuda(), torch.from_numpy(p2.astype(np.int32)).cuda()
    d_out = torch.arange(F, F + N, dtype=torch.int32, device="cuda")
    d_pool = torch.arange(0, F, dtype=torch.int32, device="cuda")
    bv_a = torch.zeros(N, dtype=torch.float64, device="cuda")
    bv_b = torch.zeros(N, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(E.gsb200_stream(ctx))

    def two_pass(i):
        dr = Draws(mode=0, seed=7, stream=i, first_unit=0)
        assert E.gsb200_cross_dev(ctx, N, d_p1.data_ptr(), d_p2.data_ptr(), 0, 0, d_out.data_ptr(), C.byref(dr)) == 0
        assert E.gsb200_gebv_dev(ctx, N, d_out.data_ptr(), 0, bv_a.data_ptr()) == 0

    def fused(i):
        dr = Draws(mode=0, seed=7, stream=i, first_unit=0)
        assert E.gsb200_cross_gebv_dev(ctx, N, d_p1.data_ptr(), d_p2.data_ptr(), 0, 0, d_out.data_ptr(), C.byref(dr),
                                       F, d_pool.data_ptr(), 0, bv_b.data_ptr()) == 0, E.gsb200_last_error()

    res = {}
    with torch.cuda.stream(stream):
        for name, fn in (("two passes", two_pass), ("fused", fused)):
            for i in range(3):
                fn(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(20):
                fn(100 + i)
            b.record(stream)
            b.synchronize()
            res[name] = a.elapsed_time(b) / 20
    assert E.gsb200_cross_dev_status(ctx) == 0, E.gsb200_last_error()
    err = float((bv_a - bv_b).abs().max().item() / (bv_a.abs().mean().item() + 1e-300))
    print("N=%d M=%d pool=%d: two passes %.3f ms, fused %.3f ms per batch (x%.2f); max rel diff %.1e" % (
        N, M, F, res["two passes"], res["fused"], res["two passes"] / res["fused"], err))


if __name__ == "__main__":
    main()
