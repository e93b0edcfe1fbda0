import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
from genomicsimulation_b200 import SimData, synthetic
M, CHR, F, N = 50000, 21, 1000, 20000
ds = synthetic.founders(F, M, CHR, 150.0, seed=2, n_effect_sets=4)
sim = SimData.from_dataset(ds)
E, ctx = sim.E, sim.ctx
assert E.gsb200_store_reserve(ctx, F + N) == 0
from genomicsimulation_b200._lib import Draws
rng = np.random.default_rng(3)
p1 = rng.integers(0, F, N).astype(np.uint32); p2 = ((p1 + rng.integers(1, F, N)) % F).astype(np.uint32)
out = np.arange(F, F + N, dtype=np.uint32)
dr = Draws(mode=0, seed=1, stream=0, first_unit=0)
assert E.gsb200_cross(ctx, N, p1.ctypes.data, p2.ctypes.data, 0, 0, out.ctypes.data, C.byref(dr)) == 0
blk = sim.blocks_evenlength(10, 1)
off, mk = blk.export()
nb = len(off) - 1
for sets in (1, 4):
    outs = [torch.zeros(2 * N * nb, dtype=torch.float64, device="cuda") for _ in range(sets)]
    ptrs = (C.c_void_p * sets)(*[o.data_ptr() for o in outs])
    effs = np.arange(sets, dtype=np.uint32)
    for rep in range(3):
        st = E.gsb200_local_gebv_multi_dev(ctx, N, out.ctypes.data, nb, off.ctypes.data, mk.ctypes.data, sets, effs.ctypes.data, ptrs)
        assert st == 0, E.gsb200_last_error()
    torch.cuda.synchronize()
print("done")
import time
stream = torch.cuda.ExternalStream(E.gsb200_stream(ctx))
for sets in (1, 4):
    outs = [torch.zeros(2 * N * nb, dtype=torch.float64, device="cuda") for _ in range(sets)]
    ptrs = (C.c_void_p * sets)(*[o.data_ptr() for o in outs])
    effs = np.arange(sets, dtype=np.uint32)
    torch.cuda.synchronize()
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(stream)
        st = E.gsb200_local_gebv_multi_dev(ctx, N, out.ctypes.data, nb, off.ctypes.data, mk.ctypes.data, sets, effs.ctypes.data, ptrs)
        b.record(stream)
        b.synchronize()
        t1 = time.perf_counter()
        print("sets %d: host %.3f ms, device span %.3f ms" % (sets, (t1 - t0) * 1e3, a.elapsed_time(b)))
