"""Regenerates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py

Outputs
  helper.json    — the reference's own 6-founder fixture (tests/testthat/helper_*.txt) as
                   loaded by the reference's loaders, plus the reference's answers for every
                   evaluation call the reference's tests pin (BVs, local BVs for three
                   blockings x two effect sets, allele counts, selection).
  scheme.json    — a 12-founder x 60-marker x 3-chromosome synthetic dataset, the draw tape the
                   reference consumed while running tests/scenarios.crossing_scheme, and the
                   resulting genotypes / ids / pedigrees / groups.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref, datasets  # noqa: E402
import scenarios  # noqa: E402

T = "/root/reference/tests/testthat/"
OUT = os.path.dirname(os.path.abspath(__file__))


def helper():
    r = ref.RefSim()
    g, m, e = r.load(T + "helper_genotypes.txt", T + "helper_map.txt", T + "helper_eff.txt")
    e2 = r.load_effects(T + "helper_eff_2.txt")
    ds = datasets.to_jsonable(datasets.from_ref(r))
    out = dict(dataset=ds, group=g)
    out["bvs"] = [r.calculate_bvs(g, e).tolist(), r.calculate_bvs(g, e2).tolist()]
    out["counts"] = {a: r.calculate_allele_counts(g, a).tolist() for a in ("T", "A")}
    blockings = {}
    for n in (3, 1):
        b = r.blocks_evenlength(n)
        off, mk = b.export()
        blockings["chrsplit%d" % n] = dict(
            offsets=off.tolist(), markers=mk.tolist(),
            local=[r.calculate_local_bvs(g, b, e).tolist(), r.calculate_local_bvs(g, b, e2).tolist()])
        b.free()
    # test-calculators.R:130-136: blocks with repeated / unknown markers, as resolved by the wrapper
    off = [0, 3, 4, 6, 9]
    mk = [0, 1, 0, 2, 0, 2, 1, 0, 2]
    b = r.blocks_from_csr(off, mk)
    blockings["custom"] = dict(offsets=off, markers=mk,
                               local=[r.calculate_local_bvs(g, b, e).tolist(), r.calculate_local_bvs(g, b, e2).tolist()])
    b.free()
    # gsc_load_blocks on the reference's own block file (test-calculators.R:104-109)
    b = r.blocks_from_file(T + "helper_blocks.txt")
    off, mk = b.export()
    blockings["file"] = dict(offsets=off.tolist(), markers=mk.tolist(), text=open(T + "helper_blocks.txt").read(),
                             local=[r.calculate_local_bvs(g, b, e).tolist(), r.calculate_local_bvs(g, b, e2).tolist()])
    b.free()
    out["blockings"] = blockings
    # centred effects (core.c:9603-9612, 10075-10076)
    r.set_centres(e, [0.25, -0.5, 0.125])
    b = r.blocks_evenlength(3)
    out["centred"] = dict(centres=[0.25, -0.5, 0.125], bvs=r.calculate_bvs(g, e).tolist(),
                          local=r.calculate_local_bvs(g, b, e).tolist())
    b.free()
    r.close()
    # selection: test-group-utils.R:206-217
    sel = []
    for top_n, low in ((5, False), (2, True), (3, False), (1, True)):
        r = ref.RefSim()
        g, m, e = r.load(T + "helper_genotypes.txt", T + "helper_map.txt", T + "helper_eff.txt")
        ng = r.split_by_bv(g, e, top_n, low)
        sel.append(dict(top_n=top_n, low=low, indexes=r.group_indexes(ng).tolist()))
        r.close()
    out["selection"] = sel
    with open(os.path.join(OUT, "helper.json"), "w") as f:
        json.dump(out, f)


def scheme():
    ds = datasets.synthetic(12, 60, 3, 100.0, seed=11, spacing="random")
    d = tempfile.mkdtemp()
    gf, mf, ef = datasets.write_files(ds, d)
    r = ref.RefSim()
    g, m, e = r.load(gf, mf, ef)
    loaded = datasets.from_ref(r)
    ref.seed(20261019)
    ref.record_start()
    groups = scenarios.crossing_scheme(r, g, small=True)
    tape = ref.record_stop()
    st = scenarios.export_state(r)
    out = dict(dataset=datasets.to_jsonable(loaded), founders=g, groups=groups,
               tape_values=[float.hex(v) for v in tape.values.tolist()], tape_kinds=tape.kinds.tolist(),
               state={k: v.tolist() for k, v in st.items()},
               bvs=r.calculate_bvs(0, e).tolist())
    with open(os.path.join(OUT, "scheme.json"), "w") as f:
        json.dump(out, f)
    r.close()


if __name__ == "__main__":
    helper()
    scheme()
    print("golden fixtures written to", OUT)
