#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config 2 (per GPU): wheat-like 21 chr x 50k SNPs,
100k random-cross offspring + GEBV of every offspring, one "step" = one such generation.

  python bench.py [--gpus N --steps K --warmup W]            this engine
  python bench.py --impl reference [...]                      the reference's own C core on the host

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "offspring genotypes/s (random cross + GEBV of every offspring)"
UNIT = "offspring/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--offspring", type=int, default=100000, help="offspring per GPU per step")
    ap.add_argument("--founders", type=int, default=1000)
    ap.add_argument("--markers", type=int, default=50000)
    ap.add_argument("--chr", type=int, default=21)
    ap.add_argument("--cpu-sample", type=int, default=16000, help="offspring in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(a):
    return {"workload": "wheat-like %d chr x %d SNPs (150 cM each), %d founders, %d random-cross offspring per GPU per step "
                        "+ GEBV of every offspring (BASELINE.json configs[1])" % (a.chr, a.markers, a.founders, a.offspring),
            "n_markers": a.markers, "n_chr": a.chr, "founders": a.founders, "offspring_per_gpu_per_step": a.offspring,
            "draws": "native Philox4x32-10 keyed by (seed, step) and indexed by offspring",
            "l2": "each step writes %.2f GB of offspring and re-reads them for GEBV (> 126 MB L2), so the next "
                  "step's parent pool is read cold" % (a.offspring * 12544 / 1e9)}


# ---------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU baselines

def _cpu_worker(args):
    """One host process: the reference's own C core (oracle/_ref) or, without it, the restatement,
    making `n` random-cross offspring + their GEBVs `steps` times."""
    kind, shape, n, steps, warm, tmpdir, seed = args
    M, Cn, founders = shape
    from oracle import datasets
    ds = datasets.synthetic(founders, M, Cn, 150.0, seed=2)
    if kind == "reference":
        from oracle import ref as backend
        gf, mf, ef = datasets.write_files(ds, tmpdir, tag="w%d" % seed)
        sim = backend.RefSim()
        backend.set_print_mode(0)
        g, m, e = sim.load(gf, mf, ef)
    else:
        from oracle import port as backend
        sim = backend.PortSim.from_dataset(ds)
        g, m, e = 1, 0, 0
    backend.seed(seed)
    times = []
    for it in range(warm + steps):
        t0 = time.perf_counter()
        grp = sim.make_random_crosses(g, n, 0, m)
        bv = sim.calculate_bvs(grp, e)
        sim.delete_group(grp)
        times.append(time.perf_counter() - t0)
        assert len(bv) == n
    return times[warm:]


def cpu_run(shape, n_total, steps, warm, procs):
    """`procs` independent host processes (the reference is single-threaded by design,
    vignettes/gSvignette.Rmd:67), each making n_total/procs offspring per step."""
    import multiprocessing as mp
    import tempfile
    from oracle import ref
    kind = "reference" if ref.available() else "port"
    per = max(1, n_total // procs)
    tmp = tempfile.mkdtemp(prefix="gsb200_cpu_")
    jobs = [(kind, shape, per, steps, warm, tmp, 100 + i) for i in range(procs)]
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    per_step = np.max(np.array(res), axis=0)        # slowest process per step
    return kind, per * procs, float(per_step.mean())


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    shape = (a.markers, a.chr, 50)
    n = max(procs, (a.cpu_sample // 4 // procs) * procs)
    kind, n_done, sec = cpu_run(shape, n, a.steps, min(a.warmup, 1), procs)
    val = n_done / sec
    cfg = workload(a)
    sample = ("%d offspring per step (%d per process x %d processes) from 50 founders of the same %d x %d shape; "
              "the per-offspring cost of the reference does not depend on the founder count"
              % (n_done, n_done // procs, procs, a.chr, a.markers))
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": min(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8 genotypes / f64 GEBV", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "allele_calls_per_s": val * 2 * a.markers}
    print(json.dumps(out))


# ---------------------------------------------------------------------------- the engine arm

def engine_arm(a):
    import torch
    import torch.distributed as dist
    from genomicsimulation_b200 import SimData, synthetic
    from genomicsimulation_b200._lib import Draws

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # the per-generation GEBV all-gather is 0.8 MB per rank: with NCCL's default protocol choice it took
        # 0.19 ms at N = 8, with LL128 0.06 ms (the 12.5 MB pool broadcast is unaffected: 0.045 ms)
        os.environ.setdefault("NCCL_PROTO", "LL128")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    M, F, n = a.markers, a.founders, a.offspring
    ds = synthetic.founders(F, M, a.chr, 150.0, seed=2)
    sim = SimData.from_dataset(ds, device=local)
    E, ctx = sim.E, sim.ctx
    sim.use_native_draws(seed=20261019)
    row_bytes = E.gsb200_store_row_bytes(ctx)
    assert E.gsb200_store_reserve(ctx, F + n) == 0, E.gsb200_last_error()
    ext = torch.cuda.ExternalStream(E.gsb200_stream(ctx), device=local)

    # ---- device-resident inputs of the kernel-level measurement
    rng = np.random.default_rng(1000 + rank)
    p1 = rng.integers(0, F, n)
    p2 = (p1 + rng.integers(1, F, n)) % F                       # a different parent, like core.c:8556-8581
    d_p1 = torch.from_numpy(p1.astype(np.int32)).cuda()
    d_p2 = torch.from_numpy(p2.astype(np.int32)).cuda()
    d_out = torch.arange(F, F + n, dtype=torch.int32, device="cuda")
    d_bv = torch.empty(n, dtype=torch.float64, device="cuda")
    # the founders occupy rows 0..F-1 of the store, contiguously: the parent pool is broadcast in place
    class _StoreView:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    pool = torch.as_tensor(_StoreView(E.gsb200_store_device_ptr(ctx), F * row_bytes), device="cuda") if world > 1 else None
    all_bv = torch.empty(world * n, dtype=torch.float64, device="cuda") if world > 1 else None
    torch.cuda.synchronize()

    def check(st):
        if st != 0:
            raise RuntimeError(E.gsb200_last_error().decode())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kernel_ms = {"cross": [], "gebv": [], "broadcast": [], "allgather": []}

    torch.cuda.synchronize()

    def step(i, timed):
        # everything is enqueued on the engine's stream (torch's current stream inside this call):
        # NCCL orders itself against it with events, so the step has no host synchronisation
        with torch.cuda.stream(ext):
            eb, e0, e1, e2, eg = ev(), ev(), ev(), ev(), ev()
            eb.record(ext)
            if world > 1:
                # generation boundary: rank 0's parent pool to everyone over NVLink, straight from
                # and into the packed stores (scattered pools go through gsb200_store_gather/scatter_rows_dev,
                # genomicsimulation_b200/sharding.py)
                dist.broadcast(pool, 0)
            dr = Draws(mode=0, seed=20261019, stream=i, first_unit=rank * n)
            e0.record(ext)
            check(E.gsb200_cross_dev(ctx, n, d_p1.data_ptr(), d_p2.data_ptr(), 0, 0, d_out.data_ptr(), C.byref(dr)))
            e1.record(ext)
            check(E.gsb200_gebv_dev(ctx, n, d_out.data_ptr(), 0, d_bv.data_ptr()))
            e2.record(ext)
            if world > 1:
                dist.all_gather_into_tensor(all_bv, d_bv)      # GEBVs for replicated selection
            eg.record(ext)
        if timed:
            kernel_ms["_ev"].append((eb, e0, e1, e2, eg))

    for i in range(a.warmup):
        step(i, False)
    check(E.gsb200_cross_dev_status(ctx))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kernel_ms["_ev"] = []
    launches0 = E.gsb200_kernel_launches(ctx)
    t0, t1 = ev(), ev()
    with ClockSampler(local) as clocks:
        t0.record(ext)
        for i in range(a.steps):
            step(a.warmup + i, True)
        t1.record(ext)
        ext.synchronize()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    check(E.gsb200_cross_dev_status(ctx))
    launches = E.gsb200_kernel_launches(ctx) - launches0
    ms_total = t0.elapsed_time(t1)
    for eb, e0, e1, e2, eg in kernel_ms["_ev"]:
        kernel_ms["broadcast"].append(eb.elapsed_time(e0))
        kernel_ms["cross"].append(e0.elapsed_time(e1))
        kernel_ms["gebv"].append(e1.elapsed_time(e2))
        kernel_ms["allgather"].append(e2.elapsed_time(eg))
    tt = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    ms_step = ms_total / a.steps
    value = world * n / (ms_step * 1e-3)

    # ---- the same step through the fused entry point: GEBVs assembled from the crossover plan and the
    #      parents' prefix tables while the offspring are spliced, instead of reading them back
    d_pool = torch.arange(0, F, dtype=torch.int32, device="cuda")
    d_bv2 = torch.empty(n, dtype=torch.float64, device="cuda")

    def fused_step(i):
        with torch.cuda.stream(ext):
            if world > 1:
                dist.broadcast(pool, 0)
            dr = Draws(mode=0, seed=20261019, stream=i, first_unit=rank * n)
            check(E.gsb200_cross_gebv_dev(ctx, n, d_p1.data_ptr(), d_p2.data_ptr(), 0, 0, d_out.data_ptr(), C.byref(dr),
                                          F, d_pool.data_ptr(), 0, d_bv2.data_ptr()))
            if world > 1:
                dist.all_gather_into_tensor(all_bv, d_bv2)

    last = a.warmup + a.steps - 1
    fused_step(last)                                     # same draws as the last timed step: same offspring, same GEBVs
    torch.cuda.synchronize()
    check(E.gsb200_cross_dev_status(ctx))
    fused_err = float((d_bv2 - d_bv).abs().max().item() / (d_bv.abs().mean().item() + 1e-300))
    for i in range(a.warmup):
        fused_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    f0, f1 = ev(), ev()
    f0.record(ext)
    for i in range(a.steps):
        fused_step(a.warmup + i)
    f1.record(ext)
    ext.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    check(E.gsb200_cross_dev_status(ctx))
    tf = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    fused_ms = float(tf.item()) / a.steps

    # ---- end to end through the reference-facing API (host buffers, copies inside the timed region)
    g_f = sim.founders
    e2e_t = []
    bv_host = np.empty(n, dtype=np.float64)
    for i in range(2 + min(a.steps, 20)):
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        grp = sim.make_random_crosses(g_f, n, 0, 1)
        ta = time.perf_counter()
        bv = sim.get_group_bvs(grp, 1, bv_host)
        tb = time.perf_counter()
        sim.delete_group(grp)
        if world > 1:
            dist.barrier()
        e2e_t.append(time.perf_counter() - t)
        if os.environ.get("GSB200_BENCH_VERBOSE") and rank == 0:
            print("e2e step: cross %.3f ms, gebv %.3f ms, delete %.3f ms" % ((ta - t) * 1e3, (tb - ta) * 1e3,
                  (time.perf_counter() - tb) * 1e3), file=sys.stderr)
        assert bv.shape[0] == n
    e2e_sec = float(np.mean(e2e_t[2:]))
    te = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sec = float(te.item())
    e2e = {"value": world * n / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": 4 * F + 4 * n, "d2h_bytes_per_step": 8 * n + 8 * n,
           "api": "gsb_make_random_crosses (pedigrees tracked) + gsb_get_group_bvs + gsb_delete_group (include/gsb200_gsc.h), "
                  "native draws: the candidate-parent rows go up, the picked pairs (for pedigrees) and the GEBVs come down"}

    # ---- roofline of the dominant kernel
    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak = float(peaks["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    cross_ms = float(np.mean(kernel_ms["cross"]))
    gebv_ms = float(np.mean(kernel_ms["gebv"]))
    alg_cross = n * (M / 2.0)                 # M/2 bytes per diploid offspring
    alg_gebv = n * (M / 4.0 + 8.0)            # M/4 bytes read + 8 written per individual
    roof = {"bound": "hbm", "kernel": "splice_flat_kernel (meiosis)", "achieved": alg_cross / (cross_ms * 1e-3) / 1e9,
            "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src, "ms_per_launch": cross_ms,
            "algorithmic_bytes_per_launch": alg_cross}
    roof["frac"] = roof["achieved"] / peak
    roof_g = {"bound": "hbm", "kernel": "gebv_counts_kernel (+ memset + gebv_finish_kernel)", "achieved": alg_gebv / (gebv_ms * 1e-3) / 1e9, "peak": peak,
              "unit": "GB/s", "traffic": None, "ms_per_launch": gebv_ms, "algorithmic_bytes_per_launch": alg_gebv}
    roof_g["frac"] = roof_g["achieved"] / peak
    # per-launch DRAM traffic from the committed ncu capture, when there is one
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        roof["traffic"] = tr.get("splice_flat_kernel")
        roof_g["traffic"] = tr.get("gebv_counts_kernel")
    except Exception:
        pass

    per_rank = None
    if world > 1:
        mine = torch.tensor([cross_ms, gebv_ms], dtype=torch.float64, device="cuda")
        everyone = torch.empty(2 * world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(everyone, mine)
        per_rank = {"cross": [round(float(x), 4) for x in everyone[0::2].tolist()], "gebv": [round(float(x), 4) for x in everyone[1::2].tolist()]}
    if rank != 0:
        return
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u1 packed genotypes / f64 GEBV", "data": "synthetic", "config": workload(a),
           "allele_calls_per_s": value * 2 * M, "meiosis_allele_calls_per_s": n * 2 * M / (cross_ms * 1e-3),
           "gebv_marker_ind_per_s": n * M / (gebv_ms * 1e-3),
           "roofline": roof, "roofline_gebv": roof_g, "e2e": e2e, "gpu_launches": int(launches),
           "clocks": clocks.summary()}
    out["fused_cross_gebv"] = {"ms_per_step": fused_ms, "value": world * n / (fused_ms * 1e-3), "unit": UNIT,
                               "max_rel_diff_vs_two_pass": fused_err,
                               "api": "gsb200_cross_gebv_dev: same offspring, GEBVs from the crossover plan + per-parent prefix "
                                      "tables (%d-row pool rebuilt every step); NOT the headline value, which keeps the two kernels" % F}
    if world > 1:
        out["per_rank_kernel_ms"] = per_rank
        out["collectives"] = {"pool_broadcast_ms": float(np.mean(kernel_ms["broadcast"])), "gebv_allgather_ms": float(np.mean(kernel_ms["allgather"])),
                              "note": "rank 0, CUDA events on the engine's stream; includes waiting for the slowest rank"}
    if world == 1 and not a.no_cpu_baseline:
        shape = (M, a.chr, 50)
        kind, n_done, sec = cpu_run(shape, a.cpu_sample, 1, 0, 1)
        out["cpu_baseline"] = {"value": n_done / sec, "unit": UNIT, "cores": 1, "kind": kind,
                               "sample": "%d random-cross offspring + GEBV from 50 founders of the same %d chr x %d SNP shape, "
                                         "1 thread (the reference is single-threaded), %.1f s" % (n_done, a.chr, M, sec)}
    print(json.dumps(out))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        engine_arm(a)


if __name__ == "__main__":
    main()
