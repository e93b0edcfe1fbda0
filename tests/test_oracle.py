"""Pins the oracles before anything is checked against them (CPU only).

1. The C restatement (oracle/gs_oracle.c) against the reference's OWN known answers,
   hard-coded here from the reference's tests with file:line, and against the committed
   fixtures in tests/golden/ that oracle/_ref (the unmodified reference) generated.
2. Where oracle/_ref is present (the build container): restatement == reference,
   byte for byte, on replayed draw tapes at larger shapes.
"""
import json
import os
import tempfile

import numpy as np
import pytest

from oracle import datasets, port, ref
import scenarios

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def helper():
    return load_json("helper.json")


def helper_sim(helper):
    return port.PortSim.from_dataset(datasets.from_jsonable(helper["dataset"]))


def test_gebv_known_answers(helper):
    """test-calculators.R:5-9 (also test-data-access.R:27)."""
    s = helper_sim(helper)
    np.testing.assert_allclose(s.calculate_bvs(1, 0), [1.4, 1.4, 1.6, -0.1, 0.6, -0.3], rtol=0, atol=1e-12)
    np.testing.assert_allclose(s.calculate_bvs(1, 1), [0.804, 0.804, 1.404, 2.502, -0.696, 1.902], rtol=0, atol=1e-12)
    # and bit-for-bit what the reference itself produced
    np.testing.assert_array_equal(s.calculate_bvs(1, 0), helper["bvs"][0])
    np.testing.assert_array_equal(s.calculate_bvs(1, 1), helper["bvs"][1])


def test_local_gebv_known_answers(helper):
    """test-calculators.R:43-76: 3 blocks per chromosome, both effect sets."""
    s = helper_sim(helper)
    off, mk = s.blocks_evenlength(3, 0, n_chr=2)
    # see.markerblocks: markers m1,m2,m3 in blocks 1,3,4 (test-calculators.R:45-47)
    block_of = np.repeat(np.arange(6), np.diff(off))
    assert list(mk) == [0, 1, 2] and list(block_of + 1) == [1, 3, 4]
    got = s.calculate_local_bvs(1, off, mk, 0)
    exp = [[0.9, 0, -0.1, -0.1, 0, 0]] * 5 + [[0.9, 0, -0.1, 0.1, 0, 0], [0.9, 0, -0.1, -0.1, 0, 0],
           [-0.8, 0, -0.1, 0.1, 0, 0], [0.9, 0, -0.5, -0.1, 0, 0], [0.9, 0, -0.5, -0.1, 0, 0],
           [-0.8, 0, -0.1, -0.1, 0, 0], [0.9, 0, -0.1, -0.1, 0, 0]]
    np.testing.assert_allclose(got, exp, rtol=0, atol=1e-12)
    got2 = s.calculate_local_bvs(1, off, mk, 1)
    assert got2[7].tolist() == [1.1, 0, 0.7, 0.3, 0, 0]          # AAA, test-calculators.R:72
    assert got2[8].tolist() == [2e-03, 0, -5e-02, -0.3, 0, 0]    # TTT, test-calculators.R:73


def test_blockings_match_reference(helper):
    """All three blockings the reference tests use (test-calculators.R:43-47, 105-109, 128-141),
    including repeated markers inside and across blocks."""
    s = helper_sim(helper)
    for name, b in helper["blockings"].items():
        if name.startswith("chrsplit"):
            off, mk = s.blocks_evenlength(int(name[-1]), 0, n_chr=2)
            assert off.tolist() == b["offsets"] and mk.tolist() == b["markers"], name
        for e in (0, 1):
            got = s.calculate_local_bvs(1, b["offsets"], b["markers"], e)
            np.testing.assert_array_equal(got, np.array(b["local"][e]), err_msg="%s eff %d" % (name, e))


def test_centred_effects(helper):
    """Centres: one subtraction of sum(centre) per genotype (core.c:9603-9612); the whole
    centre[m] per haplotype in local BVs (core.c:10075-10076)."""
    ds = datasets.from_jsonable(helper["dataset"])
    ds["effects"][0]["centre"] = np.array(helper["centred"]["centres"])
    s = port.PortSim.from_dataset(ds)
    np.testing.assert_array_equal(s.calculate_bvs(1, 0), helper["centred"]["bvs"])
    off, mk = s.blocks_evenlength(3, 0, n_chr=2)
    np.testing.assert_array_equal(s.calculate_local_bvs(1, off, mk, 0), np.array(helper["centred"]["local"]))


def test_allele_counts_known_answers(helper):
    """test-printers.R:67-111 (rows there are markers; here rows are genotypes)."""
    s = helper_sim(helper)
    t = s.calculate_allele_counts(1, "T")
    assert t.T.tolist() == [[2, 2, 2, 1, 2, 1], [0, 0, 0, 0, 2, 0], [2, 2, 1, 1, 2, 2]]
    a = s.calculate_allele_counts(0, "A")
    assert a.T.tolist() == [[0, 0, 0, 1, 0, 1], [2, 2, 2, 2, 0, 2], [0, 0, 1, 1, 0, 0]]
    np.testing.assert_array_equal(t, np.array(helper["counts"]["T"]))


def test_selection_known_answers(helper):
    """test-group-utils.R:206-217: number=5 -> 0..4; low 2 -> 3,5; 50% -> 0,1,2; low 20% -> 5."""
    for case in helper["selection"]:
        s = helper_sim(helper)
        g = s.split_by_bv(1, 0, case["top_n"], case["low"])
        assert s.group_indexes(g).tolist() == case["indexes"]
    expect = {(5, False): [0, 1, 2, 3, 4], (2, True): [3, 5], (3, False): [0, 1, 2], (1, True): [5]}
    for case in helper["selection"]:
        assert case["indexes"] == expect[(case["top_n"], case["low"])]


def test_scheme_replay_matches_reference_fixture():
    """The restatement, fed the tape the reference consumed, must rebuild the reference's
    genotypes, ids, pedigrees and groups exactly, and use up the tape exactly."""
    j = load_json("scheme.json")
    ds = datasets.from_jsonable(j["dataset"])
    tape = ref.Tape([float.fromhex(v) for v in j["tape_values"]], j["tape_kinds"])
    s = port.PortSim.from_dataset(ds)
    with port.Replay(tape) as rp:
        groups = scenarios.crossing_scheme(s, j["founders"], small=True)
    assert not rp.error and rp.consumed == len(tape)
    assert groups == j["groups"]
    st = scenarios.export_state(s)
    for k in ("ids", "ped1", "ped2", "groups"):
        assert st[k].tolist() == j["state"][k], k
    np.testing.assert_array_equal(st["genotypes"], np.array(j["state"]["genotypes"], dtype=np.uint8))
    np.testing.assert_array_equal(s.calculate_bvs(0, 0), j["bvs"])


def test_structural_invariants():
    """test-progression.R:123-230: group sizes, DH homozygosity (216-217), clone identity (229)."""
    ds = datasets.synthetic(8, 90, 3, 120.0, seed=5)
    s = port.PortSim.from_dataset(ds)
    port.seed(3)
    g2 = s.make_random_crosses(1, 7, family_size=3)
    assert s.group_size(g2) == 21
    dh = s.make_doubled_haploids(g2)
    x = s.export_group(dh)["genotypes"]
    assert np.array_equal(x[:, 0::2], x[:, 1::2])
    cl = s.make_clones(dh)
    assert np.array_equal(s.export_group(cl)["genotypes"], x)
    # offspring alleles come from the right parent strand set
    kids = s.export_group(g2)
    all_ = s.export_group(0)
    by_id = {int(i): all_["genotypes"][k] for k, i in enumerate(all_["ids"])}
    for k in range(21):
        p1, p2 = by_id[int(kids["ped1"][k])], by_id[int(kids["ped2"][k])]
        c = kids["genotypes"][k]
        assert np.all((c[0::2] == p1[0::2]) | (c[0::2] == p1[1::2]))
        assert np.all((c[1::2] == p2[0::2]) | (c[1::2] == p2[1::2]))


def test_draw_source_distributions():
    """The oracles' own sampler must be an exact Poisson / U(0,1) source, because the native
    GPU RNG is KS-tested against it."""
    from scipy import stats
    port.seed(99)
    L = port.lib()
    L.gsdraw_rpois.restype = L.gsdraw_unif.restype = __import__("ctypes").c_double
    L.gsdraw_rpois.argtypes = [__import__("ctypes").c_double]
    for mu in (0.3, 1.5, 20.0, 40.0):
        k = np.array([L.gsdraw_rpois(mu) for _ in range(20000)])
        assert abs(k.mean() - mu) < 5 * np.sqrt(mu / 20000)
        assert abs(k.var() - mu) < 0.1 * mu + 0.02
    u = np.array([L.gsdraw_unif() for _ in range(20000)])
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert L.gsdraw_rpois(0.0) == 0.0


@needs_ref
@pytest.mark.parametrize("shape", [(20, 1000, 3, 100.0), (30, 5000, 21, 150.0)])
def test_port_equals_reference_on_replay(shape):
    n, M, C, cm = shape
    ds = datasets.synthetic(n, M, C, cm, seed=M, spacing="random")
    d = tempfile.mkdtemp()
    gf, mf, ef = datasets.write_files(ds, d)
    r = ref.RefSim()
    g, m, e = r.load(gf, mf, ef)
    loaded = datasets.from_ref(r)
    np.testing.assert_array_equal(loaded["genotypes"], ds["genotypes"])
    np.testing.assert_array_equal(loaded["maps"][0]["dists"], ds["maps"][0]["dists"])
    ref.seed(M + 1)
    ref.record_start()
    groups_r = scenarios.crossing_scheme(r, g)
    tape = ref.record_stop()
    p = port.PortSim.from_dataset(loaded)
    with port.Replay(tape) as rp:
        groups_p = scenarios.crossing_scheme(p, 1)
    assert not rp.error and rp.consumed == len(tape)
    assert groups_r == groups_p
    scenarios.assert_same_state(scenarios.export_state(r), scenarios.export_state(p))
    np.testing.assert_array_equal(r.calculate_bvs(0, e), p.calculate_bvs(0, 0))
    rb = r.blocks_evenlength(4)
    off, mk = rb.export()
    off_p, mk_p = p.blocks_evenlength(4, 0, n_chr=C)
    np.testing.assert_array_equal(off, off_p)
    np.testing.assert_array_equal(mk, mk_p)
    np.testing.assert_array_equal(r.calculate_local_bvs(groups_r[0], rb, e),
                                  p.calculate_local_bvs(groups_p[0], off, mk, 0))
    np.testing.assert_array_equal(r.calculate_allele_counts(groups_r[2], "A"),
                                  p.calculate_allele_counts(groups_p[2], "A"))
    r.close()
    p.close()


@needs_ref
def test_reference_known_answers_through_shim():
    """The shimmed reference itself reproduces its tests' known answers."""
    T = "/root/reference/tests/testthat/"
    r = ref.RefSim()
    g, m, e = r.load(T + "helper_genotypes.txt", T + "helper_map.txt", T + "helper_eff.txt")
    np.testing.assert_allclose(r.calculate_bvs(g, e), [1.4, 1.4, 1.6, -0.1, 0.6, -0.3], atol=1e-12)
    b = r.blocks_evenlength(3)
    assert r.calculate_local_bvs(g, b, e)[0].tolist() == [0.9, 0, -0.1, -0.1, 0, 0]
    r.close()


@needs_ref
def test_port_names_match_reference():
    """Offspring names (prefix + counter per 1000-genotype buffer, inherited clone names) in the
    restatement against the unmodified reference."""
    ds = datasets.synthetic(9, 200, 2, 100.0, seed=61)
    d = tempfile.mkdtemp()
    gf, mf, ef = datasets.write_files(ds, d)
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    ref.seed(4)
    ref.record_start()
    groups_r = scenarios.naming_scheme(r, g)
    tape = ref.record_stop()
    p = port.PortSim.from_dataset(datasets.from_ref(r, g))
    with port.Replay(tape) as rp:
        groups_p = scenarios.naming_scheme(p, 1)
    assert not rp.error and rp.consumed == len(tape) and groups_r == groups_p
    assert r.group_names(0) == p.group_names(0)
    scenarios.assert_same_state(scenarios.export_state(r), scenarios.export_state(p))
    r.close(); p.close()
