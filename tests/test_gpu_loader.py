"""Founder genotypes from matrix files straight into the packed store (gsb_load_genotype_file,
host/gsb_simdata.c) against the UNMODIFIED reference's loader (gsc_load_genotypefile_matrix,
core.c:7085-7428) on the reference's own fixtures: allele pairs with markers as rows, slash pairs with
comma separators and no corner cell, reference-allele counts with markers as columns, a header-less
count matrix with an unknown call, and the 1003-founder file that crosses the 1000-genotype buffer.
Unphased styles draw their phase from the draw stream: the engine replays the reference's draws."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import datasets, port, ref   # noqa: E402
from test_gpu_parity import oracle_draw_pointers   # noqa: E402

pytestmark = pytest.mark.gpu
T = "/root/reference/tests/testthat/"
FIXTURES = ["helper_genotypes.txt", "helper_genotypes_slash_nocorner.txt", "helper_genotypes_inverted_counts.txt",
            "helper_genotypes_noheader_nullable_counts.txt", "helper_genotypes_long.txt"]
GOLDEN = os.path.join(ROOT, "tests", "golden", "loader")


@pytest.mark.parametrize("name", FIXTURES)
def test_loader_matches_the_reference_fixture(name):
    """golden copies of the fixtures and of what the reference loaded from them (tests/golden/make_golden_loader.py)"""
    import json
    from genomicsimulation_b200 import SimData
    js = json.load(open(os.path.join(GOLDEN, name + ".json")))
    s = SimData(len(js["marker_names"]))
    s.set_marker_names(js["marker_names"])
    s.use_host_draws(*oracle_draw_pointers())
    tape = ref.Tape([float.fromhex(v) for v in js["tape_values"]], js["tape_kinds"])
    with port.Replay(tape) as rp:
        g = s.load_genotype_file(os.path.join(GOLDEN, name))
    assert not rp.error and rp.consumed == len(tape)
    assert g == 1 and s.group_size(g) == len(js["names"])
    x = s.export_group(g)
    want = np.array(js["genotypes"], dtype=np.uint8)
    np.testing.assert_array_equal(x["genotypes"], want)
    assert s.group_names(g) == js["names"] and x["ids"].tolist() == js["ids"]
    s.close()
