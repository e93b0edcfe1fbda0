"""Parity of the CUDA engine with the oracles, through the C-ABI (GPU only).

Three-way plan of BASELINE.json's north_star:
  1. crossover replay: engine and oracle consume the SAME draw tape -> genotypes, ids,
     pedigrees, groups, allele counts must match bit for bit;
  2. GEBV / local GEBV within 1e-9 relative in fp64 (tolerance written below);
  3. the native Philox stream is KS-tested in tests/test_gpu_native.py.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import datasets, port, ref
import scenarios

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9          # north_star: "GEBVs must agree within 1e-9 relative in fp64"


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def oracle_draw_pointers():
    L = port.lib()
    return C.cast(L.gsdraw_unif, C.c_void_p), C.cast(L.gsdraw_rpois, C.c_void_p)


def engine_from(ds):
    from genomicsimulation_b200 import SimData
    s = SimData.from_dataset(ds)
    s.use_host_draws(*oracle_draw_pointers())
    return s


def bv_close(got, want, scale):
    """|got - want| <= RTOL * max(|want|, scale): `scale` = typical |BV| of the set, so that a BV
    that happens to cancel to ~0 is judged against the magnitude of what was summed."""
    got, want = np.asarray(got), np.asarray(want)
    tol = RTOL * np.maximum(np.abs(want), scale)
    bad = np.abs(got - want) > tol
    assert not bad.any(), "max abs diff %g (tol %g)" % (np.abs(got - want).max(), tol.min())


# ---------------------------------------------------------------- reference known answers

@pytest.fixture(scope="module")
def helper():
    return load_json("helper.json")


def test_helper_known_answers(helper):
    """test-calculators.R:5-9, 43-76; test-printers.R:67-111; test-group-utils.R:206-217."""
    ds = datasets.from_jsonable(helper["dataset"])
    s = engine_from(ds)
    np.testing.assert_allclose(s.calculate_bvs(1, 1), [1.4, 1.4, 1.6, -0.1, 0.6, -0.3], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(s.calculate_bvs(1, 2), [0.804, 0.804, 1.404, 2.502, -0.696, 1.902], rtol=RTOL, atol=1e-12)
    b = s.blocks_evenlength(3, 1)
    off, mk = b.export()
    assert off.tolist() == helper["blockings"]["chrsplit3"]["offsets"]
    assert mk.tolist() == helper["blockings"]["chrsplit3"]["markers"]
    for name, blk in helper["blockings"].items():
        for e in (0, 1):
            got = s.calculate_local_bvs(1, blk["offsets"], blk["markers"], e + 1)
            np.testing.assert_allclose(got, np.array(blk["local"][e]), rtol=RTOL, atol=1e-12, err_msg=name)
    t = s.calculate_allele_counts(1, "T")
    assert t.T.tolist() == [[2, 2, 2, 1, 2, 1], [0, 0, 0, 0, 2, 0], [2, 2, 1, 1, 2, 2]]
    a = s.calculate_allele_counts(0, "A")
    assert a.T.tolist() == [[0, 0, 0, 1, 0, 1], [2, 2, 2, 2, 0, 2], [0, 0, 1, 1, 0, 0]]
    x = s.export_group(0)
    np.testing.assert_array_equal(x["genotypes"], ds["genotypes"])      # pack -> unpack round trip
    s.close()
    for case in helper["selection"]:
        s = engine_from(ds)
        g = s.split_by_bv(1, 1, case["top_n"], case["low"])
        assert s.group_indexes(g).tolist() == case["indexes"]
        s.close()


def test_helper_centred(helper):
    ds = datasets.from_jsonable(helper["dataset"])
    ds["effects"][0]["centre"] = np.array(helper["centred"]["centres"])
    s = engine_from(ds)
    np.testing.assert_allclose(s.calculate_bvs(1, 1), helper["centred"]["bvs"], rtol=RTOL, atol=1e-12)
    b = s.blocks_evenlength(3, 1)
    np.testing.assert_allclose(s.calculate_local_bvs(1, b, 1), np.array(helper["centred"]["local"]), rtol=RTOL, atol=1e-12)
    s.close()


# ---------------------------------------------------------------- crossover replay

def test_scheme_replay_matches_reference_fixture():
    """The engine, fed the tape the UNMODIFIED reference consumed (tests/golden/scheme.json, made by
    tests/golden/make_golden.py), rebuilds the reference's genotypes, ids, pedigrees and groups."""
    j = load_json("scheme.json")
    ds = datasets.from_jsonable(j["dataset"])
    tape = ref.Tape([float.fromhex(v) for v in j["tape_values"]], j["tape_kinds"]).draws_only()
    s = engine_from(ds)
    with port.Replay(tape) as rp:
        groups = scenarios.crossing_scheme(s, j["founders"], small=True)
    assert not rp.error and rp.consumed == len(tape)
    assert groups == j["groups"]
    st = scenarios.export_state(s)
    for k in ("ids", "ped1", "ped2", "groups"):
        assert st[k].tolist() == j["state"][k], k
    np.testing.assert_array_equal(st["genotypes"], np.array(j["state"]["genotypes"], dtype=np.uint8))
    bv_close(s.calculate_bvs(0, 1), j["bvs"], scale=np.abs(j["bvs"]).mean())
    s.close()


def replay_both(ds, seed, small=False):
    """Oracle free-runs and records; the engine replays the same tape."""
    p = port.PortSim.from_dataset(ds)
    port.seed(seed)
    port.record_start()
    groups_p = scenarios.crossing_scheme(p, 1, small=small)
    tape = port.record_stop().draws_only()
    s = engine_from(ds)
    with port.Replay(tape) as rp:
        groups_s = scenarios.crossing_scheme(s, 1, small=small)
    assert not rp.error and rp.consumed == len(tape)
    assert groups_p == groups_s
    return p, s, groups_p


@pytest.mark.parametrize("shape", [(20, 1000, 3, 100.0, "even"), (30, 5000, 21, 150.0, "random"),
                                   (12, 50000, 21, 150.0, "random"), (16, 257, 40, 80.0, "random")])
def test_scheme_replay_matches_oracle(shape):
    n, M, Cn, cm, spacing = shape
    ds = datasets.synthetic(n, M, Cn, cm, seed=M + 7, spacing=spacing, n_effect_sets=2)
    p, s, groups = replay_both(ds, seed=M + 1)
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    want = p.calculate_bvs(0, 0)
    bv_close(s.calculate_bvs(0, 1), want, scale=np.abs(want).mean())
    want = p.calculate_bvs(groups[2], 1)
    bv_close(s.calculate_bvs(groups[2], 2), want, scale=np.abs(want).mean())
    off, mk = p.blocks_evenlength(4, 0, n_chr=Cn)
    b = s.blocks_evenlength(4, 1)
    off_s, mk_s = b.export()
    np.testing.assert_array_equal(off, off_s)
    np.testing.assert_array_equal(mk, mk_s)
    want = p.calculate_local_bvs(groups[0], off, mk, 0)
    got = s.calculate_local_bvs(groups[0], b, 1)
    assert got.shape == want.shape
    bv_close(got, want, scale=np.abs(want).mean())
    np.testing.assert_array_equal(p.calculate_allele_counts(groups[2], "A"), s.calculate_allele_counts(groups[2], "A"))
    # selection: same set, modulo ties at the cut (core.c:1305-1306)
    gp = p.split_by_bv(groups[0], 0, 5, False)
    gs = s.split_by_bv(groups[0], 1, 5, False)
    assert sorted(p.group_indexes(gp).tolist()) == sorted(s.group_indexes(gs).tolist())
    p.close()
    s.close()


def test_many_alleles_and_unknowns():
    """More than two symbols per marker and '\\0' unknowns (core.c:75) widen the store to several planes."""
    ds = datasets.synthetic(10, 700, 4, 90.0, seed=3, spacing="random", alleles=(b"A", b"C", b"G", b"T"),
                            missing_rate=0.03)
    p, s, groups = replay_both(ds, seed=11, small=True)
    assert s.E.gsb200_store_planes(s.ctx) >= 2
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    want = p.calculate_bvs(0, 0)
    bv_close(s.calculate_bvs(0, 1), want, scale=np.abs(want).mean())
    off, mk = p.blocks_evenlength(3, 0, n_chr=4)
    want = p.calculate_local_bvs(groups[1], off, mk, 0)
    bv_close(s.calculate_local_bvs(groups[1], off, mk, 1), want, scale=np.abs(want).mean() + 1e-300)
    for a in ("A", "T", "\0"):
        np.testing.assert_array_equal(p.calculate_allele_counts(0, a.encode()), s.calculate_allele_counts(0, a.encode()))
    p.close()
    s.close()


def reorder_map(ds, seed):
    """Same chromosomes, but each linkage group lists its markers through `marker_indexes`
    in an order unrelated to storage order (GSC_LINKAGEGROUP_REORDER, core.c:7911-7921)."""
    rng = np.random.default_rng(seed)
    m = ds["maps"][0]
    perm = rng.permutation(ds["n_markers"]).astype(np.uint32)
    m2 = dict(m)
    m2["marker_index"] = perm
    ds2 = dict(ds)
    ds2["maps"] = [m2]
    return ds2


def test_reorder_and_unlinked_maps():
    """The general kernel: REORDER linkage groups, and the unlinked map of core.c:6128-6152
    (one chromosome per marker, no dists, lambda 0)."""
    ds = reorder_map(datasets.synthetic(9, 400, 5, 110.0, seed=21, spacing="random"), seed=5)
    p, s, _ = replay_both(ds, seed=8, small=True)
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    p.close(); s.close()

    M = 300
    ds = datasets.synthetic(8, M, 1, 100.0, seed=22)
    ds["maps"] = [dict(lam=np.zeros(M), chr_offset=np.arange(M + 1, dtype=np.uint32), dists=np.full(M, np.nan),
                       marker_index=np.arange(M, dtype=np.uint32), has_dists=np.zeros(M, dtype=np.uint8))]
    p, s, _ = replay_both(ds, seed=9, small=True)
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    p.close(); s.close()


def test_degenerate_chromosomes():
    """One-marker chromosomes have dists = 0/0 = NaN and never switch strand (core.c:5904, 5929);
    clustered markers share a position."""
    ds = datasets.synthetic(8, 64, 1, 100.0, seed=30)
    sizes = [1, 1, 30, 1, 31]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    dists = np.zeros(64)
    rng = np.random.default_rng(2)
    for c, n in enumerate(sizes):
        lo = off[c]
        if n == 1:
            dists[lo] = np.nan
        else:
            pos = np.sort(np.round(rng.uniform(0, 50, n), 0))      # many ties
            dists[lo:lo + n] = (pos - pos[0]) / (pos[-1] - pos[0])
    ds["maps"] = [dict(lam=np.array([0.0, 0.0, 2.5, 0.0, 3.0]), chr_offset=off, dists=dists,
                       marker_index=np.arange(64, dtype=np.uint32), has_dists=np.ones(5, dtype=np.uint8))]
    p, s, _ = replay_both(ds, seed=4, small=True)
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    p.close(); s.close()


def test_empty_and_invalid_requests():
    ds = datasets.synthetic(6, 100, 2, 100.0, seed=1)
    s = engine_from(ds)
    assert s.make_random_crosses(99, 5) == 0                 # no such group
    assert s.make_random_crosses(1, 0) == 0                  # n_crosses < 1
    assert s.make_random_crosses(1, 50, 1) == 0              # cap too small
    assert s.make_targeted_crosses([100, 0xFFFFFFFF], [0, 1]) == 0   # all pairs invalid
    assert s.self_n_times(0, 1) == 0
    assert s.make_doubled_haploids(1, 7) == 0                # no such map
    assert s.calculate_bvs(1, 9).size == 0                   # no such effect set
    assert s.n_genotypes == 6
    g = s.make_random_crosses(1, 3, retain=False)            # retain = FALSE drops the offspring
    assert g == 0 and s.n_genotypes == 6
    # names and ids per 1000-genotype buffer (core.c:8168-8180)
    g = s.make_random_crosses(1, 1203, name_offspring=True, name_prefix="x")
    names = s.group_names(g)
    ids = s.export_group(g, genotypes=False)["ids"]
    # the retain=False cross above still consumed ids 7..9 (core.c:8174-8178 runs before the drop)
    assert names[0] == "x10" and names[1202] == "x1212" and ids.tolist() == list(range(10, 1213))
    s.delete_group(1)
    assert s.n_genotypes == 1203 and s.group_indexes(g).tolist() == list(range(1203))
    s.close()


def test_local_gebv_many_effect_sets_in_one_pass():
    """gsb200_local_gebv_multi (several effect sets per pass over the genotypes, the dense-contraction
    path of BASELINE.json configs[4]) against the oracle's one-set-at-a-time calls, with centres."""
    n_sets = 6
    ds = datasets.synthetic(40, 3000, 6, 130.0, seed=17, spacing="random", n_effect_sets=n_sets)
    rng = np.random.default_rng(3)
    ds["effects"][2]["centre"] = rng.standard_normal(3000) * 0.1
    p = port.PortSim.from_dataset(ds)
    s = engine_from(ds)
    off, mk = p.blocks_evenlength(7, 0, n_chr=6)
    rows = np.arange(40, dtype=np.uint32)
    nb = len(off) - 1
    outs = [np.zeros((80, nb)) for _ in range(n_sets)]
    ptrs = (C.c_void_p * n_sets)(*[o.ctypes.data for o in outs])
    effs = np.arange(n_sets, dtype=np.uint32)
    off32, mk32 = off.astype(np.uint32), mk.astype(np.uint32)
    st = s.E.gsb200_local_gebv_multi(s.ctx, 40, rows.ctypes.data, nb, off32.ctypes.data, mk32.ctypes.data,
                                     n_sets, effs.ctypes.data, ptrs)
    assert st == 0, s.E.gsb200_last_error()
    for e in range(n_sets):
        want = p.calculate_local_bvs(1, off, mk, e)
        bv_close(outs[e], want, scale=np.abs(want).mean())
        # one set at a time goes through the mma.sync kernel, six at once through the tcgen05 kernel: the same exact
        # integer digit sums, joined to fp64 in a different order — a few ulps apart
        single = s.calculate_local_bvs(1, off, mk, e + 1)
        np.testing.assert_allclose(single, outs[e], rtol=0, atol=1e-12 * np.abs(want).mean())
    p.close()
    s.close()


def test_map_that_leaves_markers_uncovered():
    """A recombination map loaded for a subset of the markers (core.c:5737-5751): gametes never write
    the uncovered markers, which keep the zeroed slot's '\\0' (core.c:75) — here the store grows a
    code for '\\0' at those markers."""
    M = 300
    ds = datasets.synthetic(8, M, 1, 100.0, seed=40)
    idx_a = np.array([m for m in range(0, 100) if not 10 <= m < 20], dtype=np.uint32)
    idx_b = np.arange(150, 300, dtype=np.uint32)
    rng = np.random.default_rng(1)
    da = np.sort(rng.uniform(0, 1, len(idx_a))); da[0], da[-1] = 0.0, 1.0
    db = np.sort(rng.uniform(0, 1, len(idx_b))); db[0], db[-1] = 0.0, 1.0
    ds["maps"] = [dict(lam=np.array([1.7, 2.2]), chr_offset=np.array([0, len(idx_a), len(idx_a) + len(idx_b)], dtype=np.uint32),
                       dists=np.concatenate([da, db]), marker_index=np.concatenate([idx_a, idx_b]),
                       has_dists=np.ones(2, dtype=np.uint8))]
    p, s, groups = replay_both(ds, seed=6, small=True)
    a, b = scenarios.export_state(p), scenarios.export_state(s)
    scenarios.assert_same_state(a, b)
    kids = b["genotypes"][8:]
    assert (kids[:, 2 * 100:2 * 150] == 0).all() and (kids[:, 20:40] == 0).all()      # uncovered -> '\0'
    want = p.calculate_bvs(0, 0)
    bv_close(s.calculate_bvs(0, 1), want, scale=np.abs(want).mean())
    np.testing.assert_array_equal(p.calculate_allele_counts(0, b"\0"), s.calculate_allele_counts(0, b"\0"))
    p.close(); s.close()


def test_change_allele_symbol():
    """gsc_change_allele_symbol (core.c:1029-1100) is a dictionary edit on the packed store."""
    ds = datasets.synthetic(6, 200, 2, 100.0, seed=41)
    s = engine_from(ds)
    assert s.E.gsb200_store_change_symbol(s.ctx, 0xFFFFFFFF, b"A", b"G") == 0
    assert s.E.gsb200_store_change_symbol(s.ctx, 7, b"T", b"C") == 0
    want = ds["genotypes"].copy()
    want[want == ord("A")] = ord("G")
    col = want[:, 14:16]
    col[col == ord("T")] = ord("C")
    np.testing.assert_array_equal(s.export_group(0)["genotypes"], want)
    # effects are keyed by symbol: 'A' effects no longer match anything, 'T' still does except at marker 7
    e = ds["effects"][0]
    eff = e["eff"].reshape(200, 2)
    isT = (want[:, 0::2] == ord("T")).astype(float) + (want[:, 1::2] == ord("T"))
    bv_close(s.calculate_bvs(0, 1), (isT * eff[:, 1]).sum(axis=1), scale=5.0)
    s.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
def test_crossing_plans_from_files_match_the_reference(tmp_path):
    """gsc_make_crosses_from_file / gsc_make_double_crosses_from_file (core.c:9248-9444) against the
    UNMODIFIED reference on a replayed tape: unknown names and short rows are skipped, double
    crosses find the F1s by pedigree."""
    ds = datasets.synthetic(8, 600, 3, 100.0, seed=44, spacing="random")
    gf, mf, ef = datasets.write_files(ds, str(tmp_path))
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    plan = tmp_path / "plan.txt"
    plan.write_text("F1\tF2\nF3\tF8\nF9\tF1\nF4\nF5\tF6\tF7\nF2\tF7\n")
    dplan = tmp_path / "dplan.txt"
    dplan.write_text("F1\tF2\tF3\tF8\nF5\tF6\tF2\tF7\nF1\tF3\tF2\tF8\nF1\tF2\tF4\tF5\n")
    ref.seed(5)
    ref.record_start()
    g1 = r.make_crosses_from_file(str(plan), m, m, family_size=2)
    g2 = r.make_double_crosses_from_file(str(dplan), m, m)
    tape = ref.record_stop().draws_only()
    loaded = datasets.from_ref(r, g)
    s = engine_from(loaded)
    with port.Replay(tape) as rp:
        h1 = s.make_crosses_from_file(str(plan), 1, 1, family_size=2)
        h2 = s.make_double_crosses_from_file(str(dplan), 1, 1)
    assert not rp.error and rp.consumed == len(tape)
    assert (g1, g2) == (h1, h2) and r.group_size(g1) == s.group_size(h1) == 8
    assert r.group_size(g2) == s.group_size(h2) and s.group_size(h2) >= 2
    scenarios.assert_same_state(scenarios.export_state(r), scenarios.export_state(s))
    r.close(); s.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
@pytest.mark.parametrize("shape", [(20, 1000, 3, 100.0), (24, 50000, 21, 150.0)])
def test_scheme_replay_matches_the_unmodified_reference(shape, tmp_path):
    """Every crossing entry point, GEBVs, local GEBVs, allele counts and selection against the
    UNMODIFIED reference core (oracle/_ref) itself: the reference loads its own text files, records
    the draws it consumes, and the engine replays them."""
    n, M, Cn, cm = shape
    ds = datasets.synthetic(n, M, Cn, cm, seed=M + 3, spacing="random")
    gf, mf, ef = datasets.write_files(ds, str(tmp_path))
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    ref.seed(M + 9)
    ref.record_start()
    groups_r = scenarios.crossing_scheme(r, g)
    tape = ref.record_stop().draws_only()
    s = engine_from(datasets.from_ref(r, g))
    with port.Replay(tape) as rp:
        groups_s = scenarios.crossing_scheme(s, 1)
    assert not rp.error and rp.consumed == len(tape)
    assert groups_r == groups_s
    scenarios.assert_same_state(scenarios.export_state(r), scenarios.export_state(s))
    want = r.calculate_bvs(0, e)
    bv_close(s.calculate_bvs(0, 1), want, scale=np.abs(want).mean())
    rb = r.blocks_evenlength(5)
    off, mk = rb.export()
    sb = s.blocks_evenlength(5, 1)
    off_s, mk_s = sb.export()
    np.testing.assert_array_equal(off, off_s)
    np.testing.assert_array_equal(mk, mk_s)
    want = r.calculate_local_bvs(groups_r[3], rb, e)
    bv_close(s.calculate_local_bvs(groups_s[3], sb, 1), want, scale=np.abs(want).mean())
    np.testing.assert_array_equal(r.calculate_allele_counts(groups_r[1], "T"), s.calculate_allele_counts(groups_s[1], "T"))
    gr = r.split_by_bv(groups_r[0], e, 7, True)
    gs = s.split_by_bv(groups_s[0], 1, 7, True)
    assert sorted(r.group_indexes(gr).tolist()) == sorted(s.group_indexes(gs).tolist())
    r.close(); s.close()


def test_gebv_tcgen05_path():
    """The experimental tcgen05/TMEM GEBV kernel (GSB200_GEBV_PATH=tc5, csrc/gebv_tc5.cu) gives the
    same GEBVs as the oracle; the path is chosen once per process, hence the subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GSB200_GEBV_PATH="tc5")
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "gebv_path_check.py")], cwd=root, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("n=")]
    assert len(lines) == 2 and all("path=tc5" in l and l.rstrip().endswith("OK") for l in lines), out.stdout


def test_config5_marker_count():
    """BASELINE.json configs[4] marker count (600 000 SNPs, 21 chr) on a handful of individuals:
    crossing replay, GEBV, local GEBV by 10 blocks per chromosome with 3 effect sets, allele counts."""
    ds = datasets.synthetic(6, 600000, 21, 180.0, seed=50, spacing="random", n_effect_sets=3)
    p = port.PortSim.from_dataset(ds)
    port.seed(77)
    port.record_start()
    gp = [p.make_random_crosses(1, 12, family_size=2), ]
    gp.append(p.self_n_times(2, gp[0]))
    gp.append(p.make_doubled_haploids(gp[1]))
    tape = port.record_stop().draws_only()
    s = engine_from(ds)
    with port.Replay(tape) as rp:
        gs = [s.make_random_crosses(1, 12, 0, 1, family_size=2), ]
        gs.append(s.self_n_times(2, gs[0], 1))
        gs.append(s.make_doubled_haploids(gs[1], 1))
    assert not rp.error and rp.consumed == len(tape) and gp == gs
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    for e in range(3):
        want = p.calculate_bvs(0, e)
        bv_close(s.calculate_bvs(0, e + 1), want, scale=np.abs(want).mean())
    off, mk = p.blocks_evenlength(10, 0, n_chr=21)
    b = s.blocks_evenlength(10, 1)
    for e in range(3):
        want = p.calculate_local_bvs(gp[2], off, mk, e)
        bv_close(s.calculate_local_bvs(gs[2], b, e + 1), want, scale=np.abs(want).mean())
    np.testing.assert_array_equal(p.calculate_allele_counts(gp[0], "A"), s.calculate_allele_counts(gs[0], "A"))
    p.close(); s.close()


def test_names_ids_and_pedigrees_match_the_oracle():
    """Offspring names, ids and pedigrees as the reference hands them out (scenarios.naming_scheme)."""
    ds = datasets.synthetic(9, 200, 2, 100.0, seed=61)
    p = port.PortSim.from_dataset(ds)
    port.seed(4)
    port.record_start()
    gp = scenarios.naming_scheme(p, 1)
    tape = port.record_stop().draws_only()
    s = engine_from(ds)
    with port.Replay(tape) as rp:
        gs = scenarios.naming_scheme(s, 1)
    assert not rp.error and rp.consumed == len(tape) and gp == gs
    assert p.group_names(0) == s.group_names(0)
    scenarios.assert_same_state(scenarios.export_state(p), scenarios.export_state(s))
    p.close(); s.close()


def test_optimal_haplotypes_known_answers(helper):
    """test-calculators.R:238-273: see.optimal.haplotype "TAA"/"AAA", see.optimal.GEBV 1.8/4.2,
    see.minimal.GEBV -2.8/-0.696, see.optimal.possible.* on the whole group and on {G05, G06}."""
    s = engine_from(datasets.from_jsonable(helper["dataset"]))
    assert s.optimal_haplotype(1) == b"TAA" and s.optimal_haplotype(2) == b"AAA"
    np.testing.assert_allclose([s.optimal_bv(1), s.minimal_bv(1), s.optimal_bv(2), s.minimal_bv(2)], [1.8, -2.8, 4.2, -0.696], atol=1e-12)
    assert s.optimal_possible_haplotype(1, 1) == b"TAA"
    np.testing.assert_allclose(s.optimal_possible_bv(1, 1), 1.8, atol=1e-12)
    g2 = s.make_group_from([4, 5])
    assert s.optimal_possible_haplotype(g2, 1) == b"TAT"
    np.testing.assert_allclose(s.optimal_possible_bv(g2, 1), 1.4, atol=1e-12)
    s.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
def test_optimal_possible_matches_the_reference(tmp_path):
    """Allele-presence OR-reduction + effect-table walk against the reference's genotype scan, on a
    4-allele dataset with unknowns (several bit-planes) and on subgroups that lack some alleles."""
    ds = datasets.synthetic(14, 900, 3, 100.0, seed=70, spacing="random", alleles=(b"A", b"C", b"G", b"T"), missing_rate=0.05)
    gf, mf, ef = datasets.write_files(ds, str(tmp_path))
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    s = engine_from(datasets.from_ref(r, g))
    for members in ([0, 1], [2, 5, 7, 11], list(range(14))):
        gr = r.make_group_from(members)
        gs = s.make_group_from(members)
        assert r.optimal_possible_haplotype(gr, e) == s.optimal_possible_haplotype(gs, 1)
        np.testing.assert_allclose(s.optimal_possible_bv(gs, 1), r.optimal_possible_bv(gr, e), rtol=1e-12)
    assert r.optimal_haplotype(e) == s.optimal_haplotype(1)
    np.testing.assert_allclose([s.optimal_bv(1), s.minimal_bv(1)], [r.optimal_bv(e), r.minimal_bv(e)], rtol=1e-12)
    r.close(); s.close()


def test_blocks_from_file(helper, tmp_path):
    """gsc_load_blocks (core.c:9850-9918): the reference's own block file (golden), then a file with
    unknown names, empty blocks and a name without its closing semicolon against the unmodified
    reference; local GEBVs on the loaded blocks.  (A blank line inside the file leaves the reference
    with an uninitialised trailing block, core.c:9855-9858 — not a case to agree on.)"""
    ds = datasets.from_jsonable(helper["dataset"])
    s = engine_from(ds)
    blk = helper["blockings"]["file"]
    f = tmp_path / "blocks.txt"
    f.write_text(blk["text"])
    b = s.blocks_from_file(str(f))
    off, mk = b.export()
    assert off.tolist() == blk["offsets"] and mk.tolist() == blk["markers"]
    for e in (0, 1):
        np.testing.assert_allclose(s.calculate_local_bvs(1, b, e + 1), np.array(blk["local"][e]), rtol=RTOL, atol=1e-12)
    b.free()
    s.close()

    ds = datasets.synthetic(10, 400, 4, 110.0, seed=77, spacing="random")
    names = ds["marker_names"]
    rng = np.random.default_rng(8)
    lines = ["Chrom\tPos\tName\tClass\tMarkers"]
    for k in range(9):
        picks = [names[i] for i in sorted(rng.choice(400, size=int(rng.integers(0, 40)), replace=False))]
        if k == 3:
            picks = picks[:5] + ["nosuchmarker", "m"] + picks[5:]
        text = "".join(p + ";" for p in picks)
        if k == 5:
            text += names[7]                                  # no closing semicolon: not a member
        lines.append("%d\t%.1f\tb%06d\tb\t%s" % (k % 4 + 1, 3.5 * k, k, text))
    f2 = tmp_path / "blocks2.txt"
    f2.write_text("\n".join(lines) + "\n")
    gf, mf, ef = datasets.write_files(ds, str(tmp_path))
    r = ref.RefSim()
    g, m, e = r.load(gf, mf, ef)
    rb = r.blocks_from_file(str(f2))
    want_off, want_mk = rb.export()
    s = engine_from(ds)
    b = s.blocks_from_file(str(f2))
    off, mk = b.export()
    assert off.tolist() == want_off.tolist() and mk.tolist() == want_mk.tolist()
    bv_close(s.calculate_local_bvs(1, b, 1), r.calculate_local_bvs(g, rb, e), scale=1.0)
    assert s.blocks_from_file(str(tmp_path / "absent.txt")).export()[0].tolist() == [0]
    rb.free(); b.free()
    r.close(); s.close()


def test_non_finite_effects_propagate_like_the_reference():
    """An effect set holding NaN / Inf has no fixed-point image: GEBVs and local GEBVs then come from the fp64
    table kernels and carry the NaN / Inf exactly where the oracle's sequential sums do (core.c:9587-9597)."""
    ds = datasets.synthetic(10, 400, 2, 100.0, seed=23)
    ds["effects"][0]["eff"][6] = np.nan          # allele 'A' of marker 3
    ds["effects"][0]["eff"][101] = np.inf        # allele 'T' of marker 50
    p = port.PortSim.from_dataset(ds)
    s = engine_from(ds)
    want, got = p.calculate_bvs(1, 0), s.calculate_bvs(1, 1)
    assert np.array_equal(np.isnan(want), np.isnan(got)) and np.array_equal(np.isposinf(want), np.isposinf(got))
    ok = np.isfinite(want)
    bv_close(got[ok], want[ok], scale=1.0)
    off, mk = p.blocks_evenlength(4, 0, n_chr=2)
    wl, gl = p.calculate_local_bvs(1, off, mk, 0), s.calculate_local_bvs(1, off, mk, 1)
    assert np.array_equal(np.isnan(wl), np.isnan(gl)) and np.array_equal(np.isinf(wl), np.isinf(gl))
    p.close()
    s.close()
