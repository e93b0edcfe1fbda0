"""The save/see boundary byte for byte: gsb_save_genotypes / _allele_counts / _pedigrees / _bvs /
_local_bvs and the save-as-you-go files of the crossing scaffold (host/gsb_simdata.c) against

  * the committed golden files tests/golden/savers/ that the UNMODIFIED reference wrote for the same
    scenario on its own 6-founder fixture (tests/golden/make_golden_savers.py; reference tests
    test-printers.R:1-65, 67-111, 200-323 and test-progression.R:54-109), replaying the draws the
    reference consumed, and the literal lines those reference tests pin;
  * the reference itself (oracle/_ref, when present) on a synthetic population whose batches cross the
    1000-genotype buffer boundary of core.c:8307-8320."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import datasets, port, ref   # noqa: E402
import scenarios   # noqa: E402
from test_gpu_parity import engine_from   # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


def same_files(want_dir, got_dir, files):
    for name in files:
        want = open(os.path.join(want_dir, name), "rb").read()
        got = open(os.path.join(got_dir, name), "rb").read()
        if want != got:
            w, g = want.split(b"\n"), got.split(b"\n")
            for k, (a, b) in enumerate(zip(w, g)):
                assert a == b, "%s line %d:\n reference %r\n engine    %r" % (name, k + 1, a[:200], b[:200])
            assert len(w) == len(g), "%s: %d lines, reference has %d" % (name, len(g), len(w))


def test_savers_match_the_reference_golden_files(tmp_path):
    js = json.load(open(os.path.join(GOLDEN, "helper.json")))
    man = json.load(open(os.path.join(GOLDEN, "savers", "manifest.json")))
    s = engine_from(datasets.from_jsonable(js["dataset"]))          # both effect sets come with the dataset
    tape = ref.Tape([float.fromhex(v) for v in man["tape_values"]], man["tape_kinds"])
    with port.Replay(tape) as rp:
        files = scenarios.saver_scheme(s, 1, (1, 2), str(tmp_path))
    assert not rp.error and rp.consumed == len(tape)
    assert files == man["files"]
    same_files(os.path.join(GOLDEN, "savers"), str(tmp_path), files)
    # the lines the reference's own tests pin
    rd = lambda n: open(os.path.join(str(tmp_path), n)).read().split("\n")
    assert rd("geno-group.txt")[:2] == ["1\tm1\tm2\tm3", "G01\tTT\tAA\tTT"]                 # test-printers.R:9-10
    assert rd("geno-all-t.txt")[:2] == ["\tG01\tG02\tG03\tG04\tG05\tG06", "m1\tTT\tTT\tTT\tTA\tTT\tAT"]   # test-printers.R:56-57
    assert rd("sayg1-genotype.txt")[:3] == ["\tm1\tm2\tm3", "G01\tTT\tAA\tTT", "G02\tTT\tAA\tTT"]       # test-progression.R:56-58
    assert rd("sayg2-bv.txt")[3].endswith("\tG04\t-0.100000")                                          # test-progression.R:73
    assert rd("sayg3-pedigree.txt")[0].endswith("\t=(G01)")                                            # test-progression.R:82
    assert rd("sayg4-bv.txt")[0] == "0\t\t1.400000" and rd("sayg4-pedigree.txt")[0] == "0\t=(G01)"     # test-progression.R:101, 109
    s.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
@pytest.mark.timeout(300)
def test_savers_match_the_reference_across_the_buffer_boundary(tmp_path):
    ds = datasets.synthetic(30, 700, 3, 100.0, seed=61, spacing="random", n_effect_sets=2)
    src = tmp_path / "in"
    gf, mf, ef = datasets.write_files(ds, str(src))
    _, _, ef2 = datasets.write_files(ds, str(src), tag="second", effect_set=1)
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    e2 = r.load_effects(ef2)
    want_dir, got_dir = tmp_path / "ref", tmp_path / "eng"
    want_dir.mkdir()
    got_dir.mkdir()
    ref.seed(99)
    ref.record_start()
    files = scenarios.saver_scheme(r, g, (e, e2), str(want_dir), big=True)       # 1200 offspring per batch
    tape = ref.record_stop().draws_only()
    s = engine_from(datasets.from_ref(r, g))
    with port.Replay(tape) as rp:
        files_s = scenarios.saver_scheme(s, 1, (1, 2), str(got_dir), big=True)
    assert not rp.error and rp.consumed == len(tape) and files == files_s
    same_files(str(want_dir), str(got_dir), files)
    assert sum(1 for _ in open(got_dir / "c1-genotype.txt")) == 1201
    scenarios.assert_same_state(scenarios.export_state(r), scenarios.export_state(s))
    r.close()
    s.close()
