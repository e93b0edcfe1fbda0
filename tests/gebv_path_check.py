"""Run by test_gpu_parity.py::test_gebv_tcgen05_path in a subprocess (the GEBV kernel is chosen once per
process by GSB200_GEBV_PATH): GEBVs of the selected path against the oracle port."""
import os, sys, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import datasets, port
from genomicsimulation_b200 import SimData
for (n, M, Cn) in ((200, 3000, 5), (1000, 50000, 21)):
    ds = datasets.synthetic(n, M, Cn, 120.0, seed=5, spacing="random")
    p = port.PortSim.from_dataset(ds)
    want = p.calculate_bvs(1, 0)
    s = SimData.from_dataset(ds)
    t = time.time(); got = s.calculate_bvs(1, 1); dt = time.time() - t
    err = np.abs(got - want).max()
    print("n=%d M=%d path=%s max abs err %.3e (scale %.2f) in %.3fs  %s" % (n, M, os.environ.get("GSB200_GEBV_PATH", "mma.sync"), err, np.abs(want).mean(), dt,
          "OK" if err <= 1e-9 * np.abs(want).mean() else "MISMATCH"))
    if err > 1e-9 * np.abs(want).mean():
        print(" first values got", got[:4], "want", want[:4])
    s.close(); p.close()
