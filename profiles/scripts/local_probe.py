"""Local GEBV timing probe: python profiles/scripts/local_probe.py N M SETS[,SETS...] [reps]
device-span (CUDA events on the engine's stream around gsb200_local_gebv_multi_dev) per number of effect
sets; GSB200_LOCAL_PATH=mma|tc5 picks the kernel (default: tc5 from 5 sets on)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from genomicsimulation_b200 import SimData, synthetic   # noqa: E402
from genomicsimulation_b200._lib import Draws           # noqa: E402

N, M = int(sys.argv[1]), int(sys.argv[2])
SETS = [int(x) for x in sys.argv[3].split(",")]
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 3
F, CHR = 64, 21
ds = synthetic.founders(F, M, CHR, 150.0, seed=2, n_effect_sets=max(SETS))
sim = SimData.from_dataset(ds)
E, ctx = sim.E, sim.ctx
assert E.gsb200_store_reserve(ctx, F + N) == 0
founders = np.arange(F, dtype=np.uint32)
dr = Draws(mode=0, seed=1, stream=0, first_unit=0)
assert E.gsb200_cross_random_pairs(ctx, N, 1, founders.ctypes.data, F, None, 0, 0, 0, None, F, C.byref(dr), None, None) == 0
blk = sim.blocks_evenlength(10, 1)
off, mk = blk.export()
nb = len(off) - 1
rows = np.arange(F, F + N, dtype=np.uint32)
stream = torch.cuda.ExternalStream(E.gsb200_stream(ctx))
row_bytes = E.gsb200_store_row_bytes(ctx)
for sets in SETS:
    outs = [torch.zeros(2 * N * nb, dtype=torch.float64, device="cuda") for _ in range(sets)]
    ptrs = (C.c_void_p * sets)(*[o.data_ptr() for o in outs])
    effs = np.arange(sets, dtype=np.uint32)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(REPS + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        st = E.gsb200_local_gebv_multi_dev(ctx, N, rows.ctypes.data, nb, off.ctypes.data, mk.ctypes.data, sets, effs.ctypes.data, ptrs)
        assert st == 0, E.gsb200_last_error()
        b.record(stream)
        b.synchronize()
        if rep:
            best = min(best, a.elapsed_time(b))
    hbm = N * row_bytes + sets * 2.0 * N * nb * 8
    print("path=%s N=%d M=%d blocks=%d sets=%2d: %8.3f ms  %.3f of HBM (6548.8 GB/s)  %.1f int8 TOPS  checksum %.6g" %
          (os.environ.get("GSB200_LOCAL_PATH", "auto"), N, M, nb, sets, best, hbm / (best * 1e-3) / 1e9 / 6548.8,
           2 * 2.0 * N * M * sets * 8 / (best * 1e-3) / 1e12, float(outs[-1][::997].sum().item())))
