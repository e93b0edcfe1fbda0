"""The chars <-> packed store boundary (K8): GB/s of chars through gsb200_store_upload_chars / _download_chars
for pageable and pinned host buffers.  python profiles/scripts/chars_boundary.py [n_genotypes] [n_markers]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from genomicsimulation_b200 import SimData, synthetic   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
ds = synthetic.founders(64, M, 21, 150.0, seed=2)
sim = SimData.from_dataset(ds)
E, ctx = sim.E, sim.ctx
assert E.gsb200_store_reserve(ctx, 64 + N) == 0
rng = np.random.default_rng(1)
chars = ds["genotypes"][rng.integers(0, 64, N)].copy()           # N x 2M, the founders' symbols
rows = np.arange(64, 64 + N, dtype=np.uint32)
pinned = torch.empty((N, 2 * M), dtype=torch.uint8).pin_memory()
pinned.numpy()[:] = chars
out_pinned = torch.empty((N, 2 * M), dtype=torch.uint8).pin_memory()
out = np.empty_like(chars)
gb = chars.nbytes / 1e9


def best(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts)


def chk(st):
    assert st == 0, E.gsb200_last_error()


for label, src in (("pageable", chars), ("pinned", pinned.numpy())):
    t = best(lambda: chk(E.gsb200_store_upload_chars(ctx, N, rows.ctypes.data, src.ctypes.data)))
    print("upload   %-8s %d x %d markers (%.2f GB of chars): %7.2f ms  %6.2f GB/s" % (label, N, M, gb, t * 1e3, gb / t))
for label, dst in (("pageable", out), ("pinned", out_pinned.numpy())):
    t = best(lambda: chk(E.gsb200_store_download_chars(ctx, N, rows.ctypes.data, dst.ctypes.data)))
    print("download %-8s %d x %d markers (%.2f GB of chars): %7.2f ms  %6.2f GB/s" % (label, N, M, gb, t * 1e3, gb / t))
    assert np.array_equal(dst, chars)
print("round trip exact")
