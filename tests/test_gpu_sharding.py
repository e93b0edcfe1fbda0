"""The sharded generation (csrc/shard.cu, csrc/select.cu) on real GPUs.

N shards must reproduce the 1-shard run exactly: selection is replicated with a deterministic tie
rule, pairs and crossovers are drawn from Philox positions indexed by the GLOBAL offspring number.
Two shards can live in one process on one GPU (their exchange buffers are then plain pointers), so
the N-ranks == 1-rank property is tested on a 1-GPU box too; the multi-process / CUDA-IPC path needs
two GPUs."""
import ctypes as C
import os
import socket
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

N0, TOP, SIZES = 64, 12, (500, 507, 514)
MAX_TOTAL, MAX_POOL = 600, 16


def dataset():
    from genomicsimulation_b200 import synthetic
    return synthetic.founders(N0, 6000, 7, 120.0, seed=4)


def reference_selection(bv, top_n, low_is_best=False):
    """ascending positions of the top_n best, ties by lower position (stable argsort)"""
    key = bv if low_is_best else -bv
    return np.sort(np.argsort(key, kind="stable")[:top_n])


def test_top_n_is_a_radix_select_with_the_reference_tie_rule():
    from genomicsimulation_b200 import SimData
    sim = SimData.from_dataset(dataset())
    rng = np.random.default_rng(3)
    cases = [rng.standard_normal(100000), np.round(rng.standard_normal(5000), 1), np.zeros(300), -np.abs(rng.standard_normal(4097)),
             np.array([1.0, 3.0, 3.0, -2.0, 0.5, 3.0, -0.0, 0.0]), rng.standard_normal(1) * 1e300, rng.standard_normal(2049) * 1e-300,
             # more ties at the cut than one block resolves for itself (select_fused_kernel's block-0 path), and a constant generation
             np.round(rng.standard_normal(200000), 0), np.full(40000, 2.5)]
    for v in cases:
        v = np.ascontiguousarray(v, dtype=np.float64)
        for top in sorted(t for t in {1, 2, len(v) // 100 + 1, len(v) // 2 + 1, len(v)} if t <= len(v)):
            for low in (0, 1):
                out = np.empty(top, dtype=np.uint32)
                st = sim.E.gsb200_top_n(sim.ctx, len(v), v.ctypes.data, top, low, out.ctypes.data)
                assert st == 0, sim.E.gsb200_last_error()
                assert out.tolist() == reference_selection(v, top, bool(low)).tolist(), (len(v), top, low)
    sim.close()


def download(sim, rows):
    g = np.zeros((len(rows), 2 * sim.M), dtype=np.uint8)
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    assert sim.E.gsb200_store_download_chars(sim.ctx, len(rows), rows.ctypes.data, g.ctypes.data) == 0
    return g


def engine_level(world):
    """`world` shards in THIS process on GPU 0, driven from one host thread through the asynchronous
    entry points: all of rank 0's generation is enqueued before rank 1's first kernel, so any host
    synchronisation inside a generation would dead-lock (and trip the 20 s exchange time-out)."""
    from genomicsimulation_b200 import SimData, sharding
    from genomicsimulation_b200._lib import Draws
    ds = dataset()
    sims = [SimData.from_dataset(ds, device=0) for _ in range(world)]
    E = sims[0].E
    for s in sims:                                             # no allocation once the generations are in flight
        assert E.gsb200_store_reserve(s.ctx, N0 + 2 * MAX_TOTAL) == 0
        rows = np.arange(N0, dtype=np.uint32)
        out = np.empty(N0)
        assert E.gsb200_gebv(s.ctx, N0, rows.ctypes.data, 0, out.ctypes.data) == 0
    shards = [sharding.Shard(sims[r], r, world, MAX_TOTAL, MAX_POOL) for r in range(world)]
    handles = b"".join(s.handle() for s in shards)
    for s in shards:
        s.connect(handles)
    n_total = N0
    spans = [sharding.offspring_range(n_total, r, world) for r in range(world)]
    d_rows = [torch.arange(lo, hi, dtype=torch.int32, device="cuda") for lo, hi in spans]   # generation 0 = the founders, sliced
    torch.cuda.synchronize()
    history = []
    for gen, n_off in enumerate(SIZES):
        ospans = [sharding.offspring_range(n_off, r, world) for r in range(world)]
        first_out = N0 + (gen % 2) * MAX_TOTAL
        for r in range(world):
            st = E.gsb200_shard_select_dev(shards[r].h, spans[r][1] - spans[r][0], d_rows[r].data_ptr(), None, n_total, 0, TOP, 0)
            assert st == 0, E.gsb200_last_error()
            dr = Draws(mode=0, seed=77, stream=gen, first_unit=ospans[r][0])
            st = E.gsb200_shard_cross_dev(shards[r].h, ospans[r][1] - ospans[r][0], 0, None, first_out, C.byref(dr), None)
            assert st == 0, E.gsb200_last_error()
        sel = [shards[r].selected(bvs=True) for r in range(world)]
        for r in range(world):
            shards[r].status()
            assert E.gsb200_cross_dev_status(sims[r].ctx) == 0
            assert sel[r][0].tolist() == sel[0][0].tolist() and np.array_equal(sel[r][2], sel[0][2])
        assert sel[0][0].tolist() == reference_selection(sel[0][2], TOP).tolist()
        geno = np.concatenate([download(sims[r], np.arange(first_out, first_out + ospans[r][1] - ospans[r][0])) for r in range(world)])
        history.append((sel[0][0].tolist(), sel[0][2].tolist(), geno))
        n_total, spans = n_off, ospans
        d_rows = [torch.arange(first_out, first_out + hi - lo, dtype=torch.int32, device="cuda") for lo, hi in spans]
        torch.cuda.synchronize()
    for s in shards:
        s.close()
    for s in sims:
        s.close()
    return history


def same_history(a, b):
    assert len(a) == len(b)
    for (sa, ba, ga), (sb, bb, gb) in zip(a, b):
        assert sa == sb and ba == bb
        assert np.array_equal(ga, gb)


@pytest.mark.timeout(300)
def test_shards_in_one_process_reproduce_one_shard():
    single = engine_level(1)
    assert len(single) == 3 and single[2][2].shape == (514, 12000)
    # selection moves the mean upward generation over generation
    assert np.mean(single[2][1]) > np.mean(single[1][1])
    same_history(engine_level(2), single)
    same_history(engine_level(3), single)


@pytest.mark.timeout(300)
def test_gametes_sorted_by_parent_change_nothing():
    """Pools that do not stay in L2 are spliced with the gametes sorted by parent (pool_random_cross, csrc/meiosis.cu);
    the draws are keyed by the global offspring number, so the offspring must be the same bit for bit.  The size
    threshold is lowered to 0 MB to take that path on this small case."""
    old = os.environ.get("GSB200_SORT_GAMETES_MIN_MB")
    try:
        # one shard first: a kernel's first launch loads its module, which waits for the device — with two shards driven from
        # one host thread that would be waiting for a kernel of shard 0 that spins until shard 1 has been enqueued
        os.environ["GSB200_SORT_GAMETES_MIN_MB"] = "0"
        sorted_one = engine_level(1)
        os.environ["GSB200_SORT_GAMETES_MIN_MB"] = "1000000"
        plain = engine_level(2)
        same_history(sorted_one, plain)
        os.environ["GSB200_SORT_GAMETES_MIN_MB"] = "0"
        same_history(engine_level(2), plain)
    finally:
        if old is None:
            os.environ.pop("GSB200_SORT_GAMETES_MIN_MB", None)
        else:
            os.environ["GSB200_SORT_GAMETES_MIN_MB"] = old


def host_level(rank, world, handles_io, device=0):
    """the host mirror's gsb_shard_* (synchronous per rank: each rank needs its own thread or process)"""
    from genomicsimulation_b200 import SimData, sharding
    ds = dataset()
    lo, hi = sharding.offspring_range(N0, rank, world)
    sim = SimData(ds["n_markers"], device=device)
    sim.add_genotypes(ds["genotypes"], ds["names"])          # every rank knows the founders (ids 1..N0) ...
    g = sim.make_group_from(np.arange(lo, hi))               # ... and evaluates its own slice of them
    sim.add_map(ds["maps"][0])
    sim.add_effects(ds["effects"][0])
    sim.use_native_draws(77)
    assert sim.E.gsb200_store_reserve(sim.ctx, N0 + 2 * MAX_TOTAL) == 0
    sim.calculate_bvs(g, 1)
    sim.shard_init(rank, world, MAX_TOTAL, MAX_POOL)
    sim.shard_connect(handles_io(rank, sim.shard_handle()))
    out = []
    n_total = N0
    for gen, n_off in enumerate(SIZES):
        new = sim.shard_select_and_cross(g, n_total, 1, TOP, n_off, name_offspring=(gen == 1), name_prefix="x")
        sel, bv = sim.shard_last_selection(n_total, TOP)
        sim.delete_group(g)
        x = sim.export_group(new)
        out.append(dict(sel=sel.tolist(), bv=bv.tolist(), geno=x["genotypes"], ids=x["ids"].tolist(), ped1=x["ped1"].tolist(),
                        ped2=x["ped2"].tolist(), names=sim.group_names(new)))
        g, n_total = new, n_off
    cur = sim.current_id
    sim.close()
    return out, cur


def merge(parts):
    """per-rank slices -> the whole generation"""
    out = []
    for gen in range(len(SIZES)):
        d = dict(sel=parts[0][gen]["sel"], bv=parts[0][gen]["bv"])
        for p in parts:
            assert p[gen]["sel"] == d["sel"] and p[gen]["bv"] == d["bv"]
        d["geno"] = np.concatenate([p[gen]["geno"] for p in parts])
        for k in ("ids", "ped1", "ped2", "names"):
            d[k] = sum((p[gen][k] for p in parts), [])
        out.append(d)
    return out


def host_level_threads(world):
    barrier = threading.Barrier(world)
    handles = [None] * world
    results = [None] * world
    errors = []

    def io(rank, h):
        handles[rank] = h
        barrier.wait(timeout=120)
        return b"".join(handles)

    def run(rank):
        try:
            results[rank] = host_level(rank, world, io)
        except BaseException as e:       # noqa: BLE001
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=280)
    assert not errors, errors
    assert len({r[1] for r in results}) == 1                   # every rank advanced current_id alike
    return merge([r[0] for r in results]), results[0][1]


@pytest.mark.timeout(400)
def test_host_mirror_shards_give_the_metadata_one_process_would():
    single, cur1 = host_level_threads(1)
    assert cur1 == N0 + sum(SIZES)
    assert single[0]["ids"] == list(range(N0 + 1, N0 + SIZES[0] + 1))
    assert single[1]["names"][:2] == ["x%d" % (N0 + SIZES[0] + 1), "x%d" % (N0 + SIZES[0] + 2)]
    # pedigrees name selected parents: generation 1's parents are ids of generation-0 offspring picked for the pool
    chosen = set(N0 + 1 + i for i in single[1]["sel"])
    assert set(single[1]["ped1"]) <= chosen and set(single[1]["ped2"]) <= chosen
    assert all(a != b for a, b in zip(single[1]["ped1"], single[1]["ped2"]))
    double, cur2 = host_level_threads(2)
    assert cur2 == cur1
    for a, b in zip(single, double):
        for k in ("sel", "bv", "ids", "ped1", "ped2", "names"):
            assert a[k] == b[k], k
        assert np.array_equal(a["geno"], b["geno"])
    # the engine-level run draws the same offspring (same seed, stream = generation, global offspring index)
    eng = engine_level(1)
    for a, e in zip(single, eng):
        assert a["sel"] == e[0] and np.array_equal(a["geno"], e[2])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genomicsimulation_b200 import sharding
    try:
        res = host_level(rank, world, lambda r, h: sharding.exchange_handles(h), device=rank)
        res[0][:] = [dict(d, geno=d["geno"].tobytes()) for d in res[0]]
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("split_push", [False, True])
def test_two_processes_two_gpus_equal_one_rank(split_push, monkeypatch):
    """one process per GPU, exchange buffers opened through CUDA IPC, handles carried by torch.distributed; also with the
    pool pushed to the peer on a second stream while the splice — gametes sorted, own parents first — waits parent by
    parent for the rows it needs (an opt-in path, GSB200_SHARD_SPLIT_PUSH, forced here)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the same property is tested in one process above")
    import torch.multiprocessing as mp
    single, cur1 = host_level_threads(1)
    monkeypatch.setenv("GSB200_SHARD_SPLIT_PUSH", "2" if split_push else "0")       # read once per (spawned) process
    if split_push:
        monkeypatch.setenv("GSB200_SORT_GAMETES_MIN_MB", "0")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=500) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = []
    for r in range(2):
        assert got[r][1] == cur1
        parts.append([dict(d, geno=np.frombuffer(d["geno"], dtype=np.uint8).reshape(-1, 12000)) for d in got[r][0]])
    double = merge(parts)
    for a, b in zip(single, double):
        for k in ("sel", "bv", "ids", "ped1", "ped2", "names"):
            assert a[k] == b[k], k
        assert np.array_equal(a["geno"], b["geno"])
