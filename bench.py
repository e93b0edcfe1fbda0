#!/usr/bin/env python
"""bench.py — BASELINE.json's metric, offspring genotypes/s, on its configs[1] shape per GPU: wheat-like
21 chr x 50k SNPs, 100k random-cross offspring + GEBV of every offspring per GPU and step.

One "step" is one generation of configs[2]'s recurrent-selection loop: GEBV of every individual of
the current generation, selection of the global top 1 % (select.by.gebv), random crosses among the
selected to make the next generation (100k offspring per GPU).  At N = 1 that is exactly configs[1]
(1 000 parents -> 100 000 offspring + their GEBVs); at N = 8 it is configs[2]'s generation at
800 000 offspring from 8 000 parents (`config3_full`, printed at N > 1, is the 1 M / 10 000 case).

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
    ap.add_argument("--cpu-sample", type=int, default=8000, help="offspring per step in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config3/4/5 extra measurements (N = 1)")
    return ap.parse_args()


def workload(a, world=1, top=None):
    n_total = world * a.offspring
    top = top if top is not None else max(2, n_total // 100)
    row_mb = a.offspring * 12544 * (a.markers / 50000.0) / 1e6
    return {"workload": "wheat-like %d chr x %d SNPs (150 cM each); per step and GPU: GEBV of 100%% of a generation's %d individuals, "
                        "top 1 %% of the whole generation (%d of %d) selected, %d random-cross offspring made from them "
                        "(BASELINE.json configs[1] shape per GPU, run as one generation of configs[2]'s loop; generation 0 = "
                        "random crosses among %d founders)" % (a.chr, a.markers, a.offspring, top, n_total, a.offspring, a.founders),
            "n_markers": a.markers, "n_chr": a.chr, "founders": a.founders, "offspring_per_gpu_per_step": a.offspring,
            "selected_parents": top, "generation_size": n_total,
            "draws": "native Philox4x32-10 keyed by (seed, generation) and indexed by global offspring number",
            "l2": "inputs larger than L2: each step writes %.0f MB of offspring per GPU and re-reads them for GEBV in the next step "
                  "(126 MB L2); the selected parents (%.1f MB) are re-read from L2 while they fit"
                  % (row_mb, top * 12544 * (a.markers / 50000.0) / 1e6)}


# ---------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md).  The sampler is started before
    the warm-up steps (nvidia-smi needs a few hundred ms to deliver its first line, more on an 8-GPU box) and every line
    is stamped on arrival; summary() keeps the lines that arrived inside the timed window, or — a timed region shorter
    than the sampling period — the ones nearest to it (the warm-up steps right before run the same load)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_lo = self.t_hi = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def wait_for_first_sample(self, seconds=3.0):
        t = time.monotonic()
        while self.proc and not self.rows and time.monotonic() - t < seconds:
            time.sleep(0.01)

    def window(self, t_lo, t_hi):
        self.t_lo, self.t_hi = t_lo, t_hi

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        rows = [r for t, r in self.rows if self.t_lo is None or self.t_lo <= t <= self.t_hi + 0.03]
        where = "inside the timed region"
        if not rows and self.rows:
            mid = 0.5 * (self.t_lo + self.t_hi)
            nearest = sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]
            rows = [r for t, r in nearest]
            where = "nearest to the timed region (%.0f ms away; the warm-up steps run the same load)" % (1e3 * min(abs(t - mid) for t, r in nearest))
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "sampled": where}


# ---------------------------------------------------------------------------- CPU baselines

def _cpu_worker(args):
    """One host process: the reference's own C core (oracle/_ref) or, without it, the restatement,
    running the same loop as the engine's step on n individuals per generation: select.by.gebv of the
    top 1 % (which computes every GEBV, core.c:9463-9517), n random crosses among the selected, delete
    the old generation."""
    kind, files, ds_args, n, steps, warm, seed = args
    if kind == "reference":
        from oracle import ref as backend
        sim = backend.RefSim()
        backend.set_print_mode(0)
        g, m, e = sim.load(*files)
    else:
        from oracle import port as backend
        from genomicsimulation_b200 import synthetic
        sim = backend.PortSim.from_dataset(synthetic.founders(*ds_args))
        g, m, e = 1, 0, 0
    backend.seed(seed)
    top = max(2, n // 100)
    cur = sim.make_random_crosses(g, n, 0, m)           # generation 0, untimed
    times = []
    for it in range(warm + steps):
        t0 = time.perf_counter()
        sel = sim.split_by_bv(cur, e, top, False)
        new = sim.make_random_crosses(sel, n, 0, m)
        sim.delete_group(cur)
        sim.delete_group(sel)
        times.append(time.perf_counter() - t0)
        cur = new
    return times[warm:]


def cpu_run(a, n_total, steps, warm, procs):
    """`procs` independent host processes (the reference is single-threaded by design,
    vignettes/gSvignette.Rmd:67), each running the loop on n_total / procs individuals per generation,
    from the SAME founders as the engine arm.  Returns (kind, offspring per step, seconds per step)."""
    import multiprocessing as mp
    import tempfile
    from oracle import datasets, ref
    from genomicsimulation_b200 import synthetic
    kind = "reference" if ref.available() else "port"
    per = max(100, n_total // procs)
    ds_args = (a.founders, a.markers, a.chr, 150.0, 2)
    files = None
    if kind == "reference":          # the reference loads text files: written once, read by every process
        tmp = tempfile.mkdtemp(prefix="gsb200_cpu_")
        files = datasets.write_files(synthetic.founders(*ds_args), tmp, tag="founders")
    jobs = [(kind, files, ds_args, per, steps, warm, 100 + i) for i in range(procs)]
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
    warm = min(a.warmup, 1)
    kind, n_done, sec = cpu_run(a, a.offspring, a.steps, warm, procs)
    val = n_done / sec
    cfg = workload(a)
    cfg["workload"] += ("; reference arm: the same loop in %d independent host processes of %d individuals per generation each "
                        "(top 1 %% selected within each process)" % (procs, n_done // procs))
    cfg["generation_size"] = n_done
    cfg["selected_parents"] = max(2, (n_done // procs) // 100) * procs
    sample = ("one full step = %d offspring (%d per process x %d processes) from the same %d founders, %d x %d shape: gsc_split_by_bv "
              "(GEBV of every individual + top 1 %%) + gsc_make_random_crosses + gsc_delete_group"
              % (n_done, n_done // procs, procs, a.founders, a.chr, a.markers))
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8 genotypes / f64 GEBV", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "allele_calls_per_s": val * 2 * a.markers}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------- the other named shapes (N = 1 extras)

def _timed(ext, fn, reps):
    """mean CUDA-event milliseconds of fn() on the engine's stream, after one untimed call"""
    import torch
    fn()
    ext.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        ext.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))


def extra_config3(a, peak):
    """configs[2] on ONE GPU: generations of 1 M offspring, top 1 % = 10 000 parents selected (the pool, 125 MB, no longer
    fits L2 beside the offspring stream) — the same sharded step as the headline, world size 1"""
    from genomicsimulation_b200 import SimData, synthetic
    M, N, F = a.markers, 1000000, a.founders
    sim = SimData.from_dataset(synthetic.founders(F, M, a.chr, 150.0, seed=2))
    sim.use_native_draws(seed=77)
    r = sharded_loop(sim, 0, 1, 0, N, F, 6, 3, 77, False)
    sim.close()
    ph = r["phases"]
    row = 12544 * (M / 50000.0)
    return {"workload": "1 000 000-offspring generations of recurrent selection on one GPU: GEBV of 1 M individuals x %d SNPs, top 1 %% = %d "
                        "parents selected, 1 M random-cross offspring (BASELINE.json configs[2], one GPU)" % (M, r["top"]),
            "regime": "pool > L2 (%.0f MB of parents against 126 MB): the gametes are spliced sorted by parent, so a parent's row comes from DRAM once and the DRAM traffic is about the M/4 written" % (r["top"] * row / 1e6),
            "ms_per_generation": r["ms_step"], "offspring_per_s": N / (r["ms_step"] * 1e-3), "phases_ms": ph,
            "splice_ms": ph["pairs_and_splice_ms"], "splice_frac_of_hbm": N * (M / 2.0) / (ph["pairs_and_splice_ms"] * 1e-3) / 1e9 / peak,
            "gebv_ms": ph["gebv_ms"], "gebv_frac_of_hbm": N * (M / 4.0 + 8) / (ph["gebv_ms"] * 1e-3) / 1e9 / peak,
            "select_ms": ph["select_ms"]}


def extra_config4(a, peak):
    """configs[3]: self.n.times to F6 (n = 5 from the F1) + make.doubled.haploids, 500 000 lines x 100 000 SNPs"""
    import torch
    from genomicsimulation_b200 import SimData, synthetic
    from genomicsimulation_b200._lib import Draws
    M, N, F = 100000, 500000, 100
    sim = SimData.from_dataset(synthetic.founders(F, M, a.chr, 150.0, seed=4))
    E, ctx = sim.E, sim.ctx
    chk = lambda st: (_ for _ in ()).throw(RuntimeError(E.gsb200_last_error().decode())) if st != 0 else None
    chk(E.gsb200_store_reserve(ctx, F + 2 * N))
    ext = torch.cuda.ExternalStream(E.gsb200_stream(ctx))
    founders = np.arange(F, dtype=np.uint32)
    dr = Draws(mode=0, seed=4, stream=0, first_unit=0)
    chk(E.gsb200_cross_random_pairs(ctx, N, 1, founders.ctypes.data, F, None, 0, 0, 0, None, F, C.byref(dr), None, None))   # the F1
    f1 = np.arange(F, F + N, dtype=np.uint32)
    f6 = np.arange(F + N, F + 2 * N, dtype=np.uint32)
    d5 = Draws(mode=0, seed=4, stream=1, first_unit=0)
    self_ms = _timed(ext, lambda: chk(E.gsb200_self_n_times(ctx, N, f1.ctypes.data, 5, 0, f6.ctypes.data, C.byref(d5))), 2)
    d6 = Draws(mode=0, seed=4, stream=2, first_unit=0)
    dh_ms = _timed(ext, lambda: chk(E.gsb200_doubled_haploid(ctx, N, f6.ctypes.data, 0, f1.ctypes.data, C.byref(d6))), 2)   # DH lines over the F1 rows
    sim.close()
    gen_bytes = N * (M / 2.0)
    return {"workload": "500 000 lines x 100 000 SNPs: self.n.times n = 5 (F1 -> F6) then make.doubled.haploids (BASELINE.json configs[3])",
            "self_n5_ms": self_ms, "doubled_haploid_ms": dh_ms, "lines_per_s": N / ((self_ms + dh_ms) * 1e-3),
            "self_frac_of_hbm_per_generation": 5 * gen_bytes / (self_ms * 1e-3) / 1e9 / peak,
            "self_frac_of_hbm_survey_8d": gen_bytes / (self_ms * 1e-3) / 1e9 / peak,
            "doubled_haploid_frac_of_hbm": N * (M / 8.0 + M / 4.0) / (dh_ms * 1e-3) / 1e9 / peak,
            "note": "per_generation counts M/2 bytes per line and generation (generations stream through HBM scratch rows, host row "
                    "lists inside the timed call); survey_8d is SURVEY 8d's M/2 per line regardless of n, which assumes the chain stays on chip"}


def extra_config5(a, peak):
    """configs[4]: local GEBVs by marker blocks, 200 000 individuals x 600 000 SNPs x 10 effect sets, 210 blocks, one pass"""
    import torch
    from genomicsimulation_b200 import SimData, synthetic
    from genomicsimulation_b200._lib import Draws
    M, N, F, S = 600000, 200000, 48, 10
    sim = SimData.from_dataset(synthetic.founders(F, M, a.chr, 150.0, seed=5, n_effect_sets=S))
    E, ctx = sim.E, sim.ctx
    chk = lambda st: (_ for _ in ()).throw(RuntimeError(E.gsb200_last_error().decode())) if st != 0 else None
    chk(E.gsb200_store_reserve(ctx, F + N))
    ext = torch.cuda.ExternalStream(E.gsb200_stream(ctx))
    founders = np.arange(F, dtype=np.uint32)
    dr = Draws(mode=0, seed=5, stream=0, first_unit=0)
    chk(E.gsb200_cross_random_pairs(ctx, N, 1, founders.ctypes.data, F, None, 0, 0, 0, None, F, C.byref(dr), None, None))
    blocks = sim.blocks_evenlength(10, 1)
    off, mk = blocks.export()
    nb = len(off) - 1
    rows = np.arange(F, F + N, dtype=np.uint32)
    effs = np.arange(S, dtype=np.uint32)
    outs = [torch.empty((2 * N, nb), dtype=torch.float64, device="cuda") for _ in range(S)]
    ptrs = (C.c_void_p * S)(*[o.data_ptr() for o in outs])
    torch.cuda.synchronize()
    call = lambda: chk(E.gsb200_local_gebv_multi_dev(ctx, N, rows.ctypes.data, nb, off.ctypes.data, mk.ctypes.data, S, effs.ctypes.data, ptrs))
    ms = _timed(ext, call, 2)
    check = float(outs[S - 1][::4001].sum().item())
    sim.close()
    row = 8 * (((M + 31) // 32 + 31) // 32 * 32)       # two strands of 128-byte aligned planes
    in_bytes, out_bytes = N * float(row), S * 2.0 * N * nb * 8
    macs = 2.0 * N * M * S * 8                   # 8 fixed-point digit columns per effect set
    return {"workload": "local GEBVs of 200 000 individuals x 600 000 SNPs x 10 effect sets over %d blocks, one pass over the genotypes, "
                        "outputs left on the device (BASELINE.json configs[4])" % nb,
            "ms": ms, "marker_haplotype_sets_per_s": 2.0 * N * M * S / (ms * 1e-3),
            "frac_of_hbm": (in_bytes + out_bytes) / (ms * 1e-3) / 1e9 / peak, "hbm_bytes": in_bytes + out_bytes,
            "int8_tensor_tops": 2 * macs / (ms * 1e-3) / 1e12, "frac_of_int8_tensor_peak": 2 * macs / (ms * 1e-3) / 1e12 / 4500.0,
            "tensor_peak_source": "nominal dense int8 4.5 POP/s (no measured int8 figure in MEASURED_PEAKS.json)", "checksum": check}


# ---------------------------------------------------------------------------- the engine arm

def phase_means(E, shard, first_epoch, last_epoch):
    """mean CUDA-event time of every phase of the generations first_epoch..last_epoch (gsb200_shard_phase_ms)"""
    acc = np.zeros(6)
    buf = (C.c_float * 6)()
    for ep in range(first_epoch, last_epoch + 1):
        if E.gsb200_shard_phase_ms(shard.h, ep, buf) != 0:
            raise RuntimeError(E.gsb200_last_error().decode())
        acc += np.array(list(buf))
    return acc / max(1, last_epoch - first_epoch + 1)


def sharded_loop(sim, rank, world, local, n, F, steps, warmup, seed, sample_clocks):
    """`steps` timed generations of the sharded recurrent-selection loop, n offspring per GPU and generation,
    through the asynchronous C-ABI entry points (device-resident row lists).  Returns timings (max over ranks)."""
    import torch
    import torch.distributed as dist
    from genomicsimulation_b200 import sharding
    from genomicsimulation_b200._lib import Draws
    E, ctx = sim.E, sim.ctx

    def check(st):
        if st != 0:
            raise RuntimeError(E.gsb200_last_error().decode())

    n_total = world * n
    top = max(2, n_total // 100)                                # select.by.gebv top 1 %
    check(E.gsb200_store_reserve(ctx, F + 2 * n))
    ext = torch.cuda.ExternalStream(E.gsb200_stream(ctx), device=local)
    shard = sharding.Shard(sim, rank, world, n_total, top)
    shard.connect_over()
    check(E.gsb200_shard_set_timing(shard.h, 1))
    # generation 0 = this rank's slice of n_total random crosses among the founders, in rows F .. F+n-1;
    # generations then alternate between the two halves of the store
    founders = np.arange(F, dtype=np.uint32)
    dr = Draws(mode=0, seed=seed, stream=0, first_unit=rank * n)
    check(E.gsb200_cross_random_pairs(ctx, n, 1, founders.ctypes.data, F, None, 0, 0, 0, None, F, C.byref(dr), None, None))
    d_rows = [torch.arange(F, F + n, dtype=torch.int32, device="cuda"), torch.arange(F + n, F + 2 * n, dtype=torch.int32, device="cuda")]
    torch.cuda.synchronize()

    def step(i):
        # ONE generation of recurrent selection, enqueued on the engine's stream without any host
        # synchronisation: GEBV of every individual of the current generation -> exchange -> replicated
        # top-1% radix selection -> exchange of the selected rows -> device pair draw -> splice
        cur = i & 1
        check(E.gsb200_shard_select_dev(shard.h, n, d_rows[cur].data_ptr(), None, n_total, 0, top, 0))
        d = Draws(mode=0, seed=seed, stream=i + 1, first_unit=rank * n)
        check(E.gsb200_shard_cross_dev(shard.h, n, 0, None, F + (1 - cur) * n, C.byref(d), None))

    ev = lambda: torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local) if sample_clocks else None
    if clocks:
        clocks.__enter__()
        clocks.wait_for_first_sample()
    for i in range(warmup):
        step(i)
    check(E.gsb200_cross_dev_status(ctx))
    shard.status()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = E.gsb200_kernel_launches(ctx)
    epoch0 = E.gsb200_shard_epoch(shard.h)
    t0, t1 = ev(), ev()
    w0 = time.monotonic()
    t0.record(ext)
    for i in range(steps):
        step(warmup + i)
    t1.record(ext)
    ext.synchronize()
    torch.cuda.synchronize()
    if clocks:
        clocks.window(w0, time.monotonic())
        clocks.__exit__()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    check(E.gsb200_cross_dev_status(ctx))
    shard.status()
    launches = E.gsb200_kernel_launches(ctx) - launches0
    timed = min(steps, 250)                                     # the event ring keeps the last 256 generations
    ph = phase_means(E, shard, epoch0 + steps - timed + 1, epoch0 + steps)
    tt = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        mine = torch.tensor(list(ph), dtype=torch.float64, device="cuda")
        everyone = torch.empty(6 * world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(everyone, mine)
        ev_ = everyone.view(world, 6)
        per_rank = {"gebv": [round(float(x), 4) for x in ev_[:, 0].tolist()], "splice": [round(float(x), 4) for x in ev_[:, 5].tolist()]}
    shard.close()
    phases = {"gebv_ms": float(ph[0]), "gebv_exchange_ms": float(ph[1]), "select_ms": float(ph[2]), "pool_push_ms": float(ph[3]),
              "pool_wait_ms": float(ph[4]), "pairs_and_splice_ms": float(ph[5])}
    return {"ms_step": float(tt.item()) / steps, "phases": phases, "launches": int(launches), "per_rank": per_rank,
            "clocks": clocks.summary() if clocks else None, "n_total": n_total, "top": top}


def engine_arm(a):
    import torch
    import torch.distributed as dist
    from genomicsimulation_b200 import SimData, sharding, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # torch.distributed is plumbing only: it carries the shards' 128-byte exchange handles once, the
        # barriers around the timed region and the max-over-ranks of the timings.  The data path of a
        # generation is the engine's own P2P push kernels (csrc/shard.cu).
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    M, F, n = a.markers, a.founders, a.offspring
    seed = 20261019
    ds = synthetic.founders(F, M, a.chr, 150.0, seed=2)
    sim = SimData.from_dataset(ds, device=local)
    sim.use_native_draws(seed=seed)
    res = sharded_loop(sim, rank, world, local, n, F, a.steps, a.warmup, seed, True)
    n_total, top, ms_step, phases = res["n_total"], res["top"], res["ms_step"], res["phases"]
    value = n_total / (ms_step * 1e-3)
    gebv_ms, cross_ms = phases["gebv_ms"], phases["pairs_and_splice_ms"]
    launches, per_rank, clocks = res["launches"], res["per_rank"], res["clocks"]

    # configs[2] at its named size when the ranks are there for it: 1 M offspring per generation from 10 000 parents
    config3_full = None
    if world > 1:
        n3 = 1000000 // world
        r3 = sharded_loop(sim, rank, world, local, n3, F, min(a.steps, 20), 3, seed + 1, False)
        config3_full = {"workload": "1 000 000-offspring generations of recurrent selection (top 1 %% = %d parents) sharded over %d GPUs, "
                                    "%d per GPU (BASELINE.json configs[2])" % (r3["top"], world, n3),
                        "ms_per_generation": r3["ms_step"], "value": r3["n_total"] / (r3["ms_step"] * 1e-3), "unit": UNIT,
                        "phases_ms": r3["phases"]}

    # ---- end to end through the reference-facing host API (host buffers, copies inside the timed region)
    g = sim.make_random_crosses(sim.founders, n, 0, 1)          # this rank's slice of a generation, as a group of the host mirror
    sim.shard_init(rank, world, n_total, top)
    sim.shard_connect(sharding.exchange_handles(sim.shard_handle()))
    bv_host = np.empty(n, dtype=np.float64)
    E = sim.E
    # the timed steps are bracketed ONCE by a barrier + synchronize on both sides; inside, the ranks meet only through the
    # generation's own exchange (two barriers per step cost more than the step at N = 8)
    e2e_warm, e2e_steps = 2, min(a.steps, 20)
    t_start = 0.0
    for i in range(e2e_warm + e2e_steps):
        if i == e2e_warm:
            sim.shard_finish()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t_start = time.perf_counter()
        t = time.perf_counter()
        new = sim.shard_select_and_cross(g, n_total, 1, top, n_total)      # pedigrees tracked, ids allocated
        ta = time.perf_counter()
        bv = sim.shard_local_bvs(bv_host)
        tb = time.perf_counter()
        sim.delete_group(g)
        g = new
        if os.environ.get("GSB200_BENCH_VERBOSE") and rank == 0:
            print("e2e step %d: select_and_cross %.3f ms, local_bvs %.3f ms, delete_group %.3f ms" %
                  (i, (ta - t) * 1e3, (tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3), file=sys.stderr)
        assert bv.shape[0] == n
    sim.shard_finish()                                                      # the last generation's offspring are in place
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t_start) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sec = float(te.item())
    e2e = {"value": n_total / e2e_sec, "unit": UNIT, "ms_per_step": e2e_sec * 1e3,
           "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 16 * n + 4 * top,
           "api": "gsb_shard_select_and_cross (pedigrees tracked, ids allocated) + gsb_shard_get_local_bvs + gsb_delete_group "
                  "(include/gsb200_gsc.h), native draws: the generation's ids go up (its rows and the offspring's are ascending runs of "
                  "the store in this loop, so only their first row numbers go with the calls; row lists of 4 bytes per genotype "
                  "would go up otherwise); the picked pairs, the selected parents' ids (for pedigrees) and the GEBVs come down"}

    # ---- the same work through the calls a user of the reference makes on ONE GPU (round 1's e2e): gsc_make_random_crosses from
    #      the founders + gsc_get_group_bvs of the offspring + gsc_delete_group, host buffers, pedigrees tracked
    e2e_api = None
    if world == 1:
        sim.shard_finish()
        sim.delete_group(g)
        k_warm, k_steps = 2, min(a.steps, 20)
        t0 = 0.0
        for i in range(k_warm + k_steps):
            if i == k_warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            new = sim.make_random_crosses(sim.founders, n, 0, 1, track_pedigree=True)
            bv = sim.get_group_bvs(new, 1, out=bv_host)
            sim.delete_group(new)
            assert bv.shape[0] == n
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / k_steps
        e2e_api = {"value": n / sec, "unit": UNIT, "ms_per_step": sec * 1e3,
                   "api": "gsb_make_random_crosses (%d founders -> %d offspring, pedigrees tracked) + gsb_get_group_bvs + gsb_delete_group, "
                          "native draws, host buffers: no selection step (round 1's e2e definition)" % (F, n)}

    # ---- rooflines of the two dominant kernels
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    alg_cross = n * (M / 2.0)                 # M/2 bytes per diploid offspring
    alg_gebv = n * (M / 4.0 + 8.0)            # M/4 bytes read + 8 written per individual
    pool_mb = top * 12544 * (M / 50000.0) / 1e6
    roof = {"bound": "hbm", "kernel": "splice_flat_kernel (meiosis; the pair-draw kernel, ~3 us, runs beside the GEBV kernel on a second stream)",
            "achieved": alg_cross / (cross_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src,
            "ms_per_launch": cross_ms, "algorithmic_bytes_per_launch": alg_cross,
            "regime": ("parent pool of %d rows = %.1f MB: re-read from L2" % (top, pool_mb)) if pool_mb < 24 else
                      ("parent pool of %d rows = %.1f MB does not stay in L2 beside the offspring stream: gametes spliced sorted by parent "
                       "(each parent's row comes from DRAM once; the phase includes the sort)" % (top, pool_mb))}
    roof["frac"] = roof["achieved"] / peak
    # the GEBV kernel is bound by instruction issue, not by HBM (alu + imma pipes busy 99.9 % of the active
    # cycles in the committed ncu capture); achieved / peak / frac stay the HBM figures the north star asks for
    roof_g = {"bound": "issue", "kernel": "gebv_counts_kernel (+ memset + gebv_finish_kernel)", "achieved": alg_gebv / (gebv_ms * 1e-3) / 1e9,
              "peak": peak, "unit": "GB/s", "traffic": None, "ms_per_launch": gebv_ms, "algorithmic_bytes_per_launch": alg_gebv}
    roof_g["frac"] = roof_g["achieved"] / peak
    try:        # per-launch DRAM traffic from the committed ncu capture of this same command (profiles/)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        roof["traffic"] = tr.get("splice_flat_kernel")
        roof_g["traffic"] = tr.get("gebv_counts_kernel")
        roof["traffic_source"] = roof_g["traffic_source"] = tr.get("source")
        if roof["traffic"] and world == 1 and n == 100000 and M == 50000:
            # the same kernel in DRAM terms: with the pool in L2 it is a pure write stream, and a pure write stream reaches
            # 7.4 TB/s on this GPU (profiles/r2_write_bw_probe.txt) — the kernel is bound by its instructions, not by HBM
            wr_peak = 7446.0
            roof["dram_terms"] = {"achieved": roof["traffic"] / (cross_ms * 1e-3) / 1e9, "peak": wr_peak, "unit": "GB/s",
                                  "frac": roof["traffic"] / (cross_ms * 1e-3) / 1e9 / wr_peak,
                                  "peak_source": "write-only stream (torch zero_ of 4 GiB), profiles/r2_write_bw_probe.txt",
                                  "bound": "issue / latency: 1 270 instructions per gamete, 56.8 % issue-active at 40 warps per SM "
                                           "(profiles/r2_ncu_full_summary.txt; DESIGN.md 7.2)"}
        if tr.get("gebv_issue"):
            gi = dict(tr["gebv_issue"])
            gi["frac"] = gi["alu_pipe_active_frac"] + gi["imma_pipe_active_frac"]
            roof_g["issue"] = gi
    except Exception:
        pass

    extras = {}
    if world == 1 and not a.no_extras:
        sim.close()
        sim = None
        torch.cuda.empty_cache()
        for name, fn in (("config3_1gpu", extra_config3), ("config4", extra_config4), ("config5", extra_config5)):
            try:
                extras[name] = fn(a, peak)
            except Exception as e:           # an extra must never cost the headline line
                extras[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        cfg = workload(a, world, top)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u1 packed genotypes / f64 GEBV", "data": "synthetic", "config": cfg,
               "allele_calls_per_s": value * 2 * M, "meiosis_allele_calls_per_s": n * 2 * M / (cross_ms * 1e-3),
               "gebv_marker_ind_per_s": n * M / (gebv_ms * 1e-3),
               "roofline": roof, "roofline_gebv": roof_g, "e2e": e2e, "gpu_launches": int(launches),
               "clocks": clocks, "phases_ms": phases,
               "collectives": {"gebv_exchange_ms": phases["gebv_exchange_ms"], "pool_push_ms": phases["pool_push_ms"],
                               "pool_wait_ms": phases["pool_wait_ms"],
                               "note": "rank 0, CUDA events on the engine's stream; P2P push kernels over NVLink (no NCCL on the data "
                                       "path); the exchange and wait phases include waiting for the slowest rank"},
               "select_ms": phases["select_ms"]}
        if per_rank:
            out["per_rank_kernel_ms"] = per_rank
        if config3_full:
            out["config3_full"] = config3_full
        if e2e_api:
            out["e2e_single_gpu_api"] = e2e_api
        out.update(extras)
        if world == 1 and not a.no_cpu_baseline:
            kind, n_done, sec = cpu_run(a, a.cpu_sample, 1, 0, 1)
            out["cpu_baseline"] = {"value": n_done / sec, "unit": UNIT, "cores": 1, "kind": kind,
                                   "sample": "one step of the same loop on %d individuals per generation (gsc_split_by_bv top 1 %% + "
                                             "gsc_make_random_crosses + gsc_delete_group) from the same %d founders, %d chr x %d SNPs, "
                                             "1 thread (the reference is single-threaded), %.1f s" % (n_done, a.founders, a.chr, M, sec)}
        print(json.dumps(out), flush=True)
    if sim is not None:
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
