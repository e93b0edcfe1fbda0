#!/usr/bin/env python
"""Times gsb200_gebv_dev alone (CUDA events, rows resident) on the wheat shape for the path chosen
by GSB200_GEBV_PATH and prints a digest of the result:  python profiles/scripts/gebv_ab.py [N] [M]"""
import ctypes as C
import hashlib
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
    F = 1000
    ds = synthetic.founders(F, M, 21, 150.0, seed=2, n_effect_sets=1)
    sim = SimData.from_dataset(ds)
    E, ctx = sim.E, sim.ctx
    assert E.gsb200_store_reserve(ctx, F + N) == 0
    rng = np.random.default_rng(3)
    p1 = rng.integers(0, F, N).astype(np.uint32)
    p2 = ((p1 + rng.integers(1, F, N)) % F).astype(np.uint32)
    out = np.arange(F, F + N, dtype=np.uint32)
    dr = Draws(mode=0, seed=1, stream=0, first_unit=0)
    assert E.gsb200_cross(ctx, N, p1.ctypes.data, p2.ctypes.data, 0, 0, out.ctypes.data, C.byref(dr)) == 0
    stream = torch.cuda.ExternalStream(E.gsb200_stream(ctx))
    d_rows = torch.from_numpy(out.astype(np.int32)).cuda()
    d_bv = torch.zeros(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ts = []
    with torch.cuda.stream(stream):
        for it in range(25):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            st = E.gsb200_gebv_dev(ctx, N, d_rows.data_ptr(), 0, d_bv.data_ptr())
            b.record(stream)
            assert st == 0, E.gsb200_last_error()
            b.synchronize()
            ts.append(a.elapsed_time(b))
    # ... and back to back without a host synchronisation in between (what the kernel does inside a long step)
    with torch.cuda.stream(stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for it in range(300):
            assert E.gsb200_gebv_dev(ctx, N, d_rows.data_ptr(), 0, d_bv.data_ptr()) == 0
        b.record(stream)
        b.synchronize()
    print("sustained: 300 calls back to back, %.4f ms per call" % (a.elapsed_time(b) / 300))
    # ... and each call behind a kernel that leaves L2 full of dirty lines (what the splice does in the step): 256 MB written
    scratch = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
    for label, dirty in (("clean", False), ("behind 256 MB of writes", True), ("clean", False), ("behind 256 MB of writes", True)):
        evs = []
        with torch.cuda.stream(stream):
            for it in range(100):
                if dirty:
                    scratch.fill_(it)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                assert E.gsb200_gebv_dev(ctx, N, d_rows.data_ptr(), 0, d_bv.data_ptr()) == 0
                b.record(stream)
                evs.append((a, b))
            stream.synchronize()
        print("back to back, %-24s %.4f ms per call" % (label + ":", sum(a.elapsed_time(b) for a, b in evs[10:]) / 90))
    ts = sorted(ts[5:])
    bv = d_bv.cpu().numpy()
    alg = N * (M / 4 + 8)
    med = ts[len(ts) // 2]
    ref_path = "/tmp/gebv_ab_ref_%d_%d.npy" % (N, M)
    if os.path.exists(ref_path):
        ref = np.load(ref_path)
        err = float(np.max(np.abs(bv - ref) / np.maximum(np.abs(ref), 1e-300)))
    else:
        np.save(ref_path, bv)
        err = 0.0
    print("path=%-8s N=%d M=%d  median %.4f ms  min %.4f ms  %.1f GB/s alg  frac %.3f  digest %s  max rel diff vs first path run %.2e" % (
        os.environ.get("GSB200_GEBV_PATH", "default"), N, M, med, ts[0], alg / med / 1e6, alg / med / 1e6 / 6548.8,
        hashlib.sha1(bv.tobytes()).hexdigest()[:12], err))


if __name__ == "__main__":
    main()
