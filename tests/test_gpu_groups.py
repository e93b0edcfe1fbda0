"""Group-split utilities of the host mirror (gsb_split_*) against the UNMODIFIED
reference (gsc_split_into_families / _halfsib_families / _individuals, gsc_split_evenly_into_two / _into_n /
_into_buckets, gsc_split_randomly_into_two / _into_n, gsc_split_by_probabilities, core.c:2980-3784) on a
replayed draw tape: the same new group numbers, in the same order, and the same group of every genotype."""
import numpy as np
import pytest

from oracle import datasets, port, ref
import scenarios
from test_gpu_parity import engine_from

pytestmark = pytest.mark.gpu


def both(tmp_path, n_founders=10, family_size=3):
    """the reference and the engine holding the same founders and the same 3-member full-sib families"""
    ds = datasets.synthetic(n_founders, 300, 2, 100.0, seed=91, spacing="random")
    gf, mf, ef = datasets.write_files(ds, str(tmp_path))
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    p1 = np.array([0, 0, 1, 2, 3, 0, 4, 2], dtype=np.uint32)
    p2 = np.array([1, 2, 2, 5, 4, 1, 3, 6], dtype=np.uint32)          # (0, 1) twice: one family; (3, 4) and (4, 3): two
    ref.seed(17)
    ref.record_start()
    f1 = r.make_targeted_crosses(p1, p2, m, m, family_size=family_size)
    tape = ref.record_stop().draws_only()
    s = engine_from(datasets.from_ref(r, g))
    with port.Replay(tape) as rp:
        h1 = s.make_targeted_crosses(p1, p2, 1, 1, family_size=family_size)
    assert not rp.error and rp.consumed == len(tape) and f1 == h1
    return r, s, f1


def run_both(r, s, kind, group, engine_call, **kw):
    ref.seed(23 + kind)
    ref.record_start()
    want = r.split(kind, group, **kw)
    tape = ref.record_stop().draws_only()
    with port.Replay(tape) as rp:
        got = engine_call()
    assert not rp.error and rp.consumed == len(tape), (kind, rp.consumed, len(tape))
    assert got == want, (kind, got, want)
    a, b = scenarios.export_state(r), scenarios.export_state(s)
    np.testing.assert_array_equal(a["groups"], b["groups"])
    return got


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
def test_family_splits_match_the_reference(tmp_path):
    r, s, f1 = both(tmp_path)
    fams = run_both(r, s, 0, f1, lambda: s.split_into_families(f1))
    assert len(fams) == 7 and sorted(s.group_size(g) for g in fams) == [3, 3, 3, 3, 3, 3, 6]
    merged = s.combine_groups(fams)
    assert r.combine_groups(fams) == merged
    halves = run_both(r, s, 1, merged, lambda: s.split_into_halfsib_families(merged, 1))
    assert len(halves) == 5                                           # first parents 0, 1, 2, 3, 4
    merged = s.combine_groups(halves)
    assert r.combine_groups(halves) == merged
    run_both(r, s, 2, merged, lambda: s.split_into_halfsib_families(merged, 2))
    # the founders (no parents) are one family; individuals of a small group get a group each
    run_both(r, s, 0, 1, lambda: s.split_into_families(1))
    some = s.make_group_from([0, 3, 4])
    assert r.make_group_from([0, 3, 4]) == some
    ones = run_both(r, s, 3, some, lambda: s.split_into_individuals(some))
    assert len(ones) == 3 and all(s.group_size(g) == 1 for g in ones)
    assert s.split_into_halfsib_families(merged, 3) == []             # core.c:3130-3134
    r.close(); s.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
def test_allocation_splits_match_the_reference(tmp_path):
    r, s, f1 = both(tmp_path, family_size=4)                           # 32 offspring
    g2 = run_both(r, s, 4, f1, lambda: [s.split_evenly_into_two(f1)])
    assert s.group_size(g2[0]) == 16 and s.group_size(f1) == 16
    thirds = run_both(r, s, 5, f1, lambda: s.split_evenly_into_n(f1, 3), n=3)
    assert sorted(s.group_size(g) for g in thirds) == [5, 5, 6] and s.group_size(f1) == 0
    buckets = run_both(r, s, 6, g2[0], lambda: s.split_into_buckets(g2[0], [3, 7]), n=3, counts=[3, 7])
    assert [s.group_size(g) for g in buckets] == [3, 7, 6]
    pool = s.combine_groups(thirds + buckets)
    assert r.combine_groups(thirds + buckets) == pool and s.group_size(pool) == 32
    two = run_both(r, s, 7, pool, lambda: [s.split_randomly_into_two(pool)])
    assert s.group_size(two[0]) + s.group_size(pool) == 32
    run_both(r, s, 8, pool, lambda: s.split_randomly_into_n(pool, 4), n=4)
    run_both(r, s, 9, two[0], lambda: s.split_by_probabilities(two[0], [0.2, 0.5]), n=3, probs=[0.2, 0.5])
    # probabilities that leave some members where they are (core.c:3388-3392), n = 1 refused (core.c:3448-3451)
    run_both(r, s, 9, 1, lambda: s.split_by_probabilities(1, [0.3]), n=2, probs=[0.3])
    assert s.split_evenly_into_n(1, 1) == [] and s.split_evenly_into_two(999) == 0
    r.close(); s.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not present")
@pytest.mark.parametrize("window,certain", [(1, True), (1, False), (3, True), (5, False)])
def test_find_crossovers_file_matches_the_reference(tmp_path, window, certain):
    """gsc_calculate_recombinations_from_file (core.c:7760-7818; find.crossovers): the table the
    reference writes for named (offspring, parent 1, parent 2) triples, byte for byte — its chromosome-local indexing
    and the `previous` that carries over between chromosomes included."""
    ds = datasets.synthetic(8, 240, 3, 100.0, seed=5, spacing="random")
    gf, mf, ef = datasets.write_files(ds, str(tmp_path))
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(gf, mf, ef)
    p1 = np.array([0, 2, 4, 1], dtype=np.uint32)
    p2 = np.array([1, 3, 5, 6], dtype=np.uint32)
    ref.seed(3)
    ref.record_start()
    f1 = r.make_targeted_crosses(p1, p2, m, m, family_size=2, name_offspring=True, name_prefix="x")
    tape = ref.record_stop().draws_only()
    s = engine_from(datasets.from_ref(r, g))
    with port.Replay(tape) as rp:
        h1 = s.make_targeted_crosses(p1, p2, 1, 1, family_size=2, name_offspring=True, name_prefix="x")
    assert not rp.error and rp.consumed == len(tape) and f1 == h1
    kids, founders = r.group_names(f1), r.group_names(g)
    assert kids == s.group_names(h1) and len(kids) == 8
    plan = tmp_path / "triples.txt"
    lines = ["%s\t%s\t%s" % (kids[k], founders[int(p1[k // 2])], founders[int(p2[k // 2])]) for k in range(8)]
    lines.insert(3, "nobody\t%s\t%s" % (founders[0], founders[1]))            # a name that does not exist: skipped with a NOTE
    plan.write_text("\n".join(lines) + "\n")
    want, got = tmp_path / "want.txt", tmp_path / "got.txt"
    assert r.find_crossovers(plan, want, window, certain) == 0
    assert s.find_crossovers(plan, got, window, certain) == 0
    a, b = want.read_bytes(), got.read_bytes()
    assert a == b and a.count(b"\n") == 9 and len(a) > 3000
    # the array form, for one triple
    o = s.min_recombinations(int(p1[0]), 101, int(p2[0]), 202, s.group_indexes(h1)[0], window, certain)
    assert set(np.unique(o).tolist()) <= {0, 101, 202} and (o != 0).any()
    r.close(); s.close()
