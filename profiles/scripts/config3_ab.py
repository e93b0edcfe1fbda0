"""config3_1gpu (bench.py extra) with and without the L2 set-aside for the parent pool: GSB200_L2_PERSIST=0|1 python profiles/scripts/config3_ab.py"""
import json, os, sys, types
sys.path.insert(0, os.getcwd())
import bench
a = types.SimpleNamespace(markers=50000, chr=21, founders=1000, offspring=100000)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6548.8
r = bench.extra_config3(a, peak)
print("L2_PERSIST=%s" % os.environ.get("GSB200_L2_PERSIST", "default"), json.dumps({k: r[k] for k in ("ms_per_generation", "splice_ms", "splice_frac_of_hbm", "gebv_ms", "gebv_frac_of_hbm", "select_ms", "phases_ms")}))
