"""The R boundary, compiled: 24 hot-path functions of the reference's UNMODIFIED src/r-wrappers.c
(SXP_make_random_crosses, SXP_self_n_times, SXP_break_group_by_GEBV_num, SXP_see_local_gebvs,
SXP_save_genotypes, ...; list in oracle/extract_wrappers.py) are extracted at build time from where
the reference lies, compiled against include/gsb200_compat.h — the reference's own names and struct
types over the device-backed mirror — and a minimal Rinternals (oracle/shim_sexp), and called here
with SEXP arguments exactly as R's .Call would (genomicSimulation-init.c:6-69)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import datasets   # noqa: E402

pytestmark = pytest.mark.gpu
WRAP = os.path.join(ROOT, "oracle", "_ref", "libgswrap.so")
VP = C.c_void_p


@pytest.fixture(scope="module")
def W():
    if not os.path.exists(WRAP):
        pytest.skip("oracle/_ref/libgswrap.so not built (make -C oracle wrappers, needs /root/reference)")
    from genomicsimulation_b200 import _lib
    _lib.host()
    L = C.CDLL(WRAP)
    for name in ("shim_int_vector", "shim_real_vector", "shim_logical", "shim_string", "shim_nil", "shim_extptr", "shim_data",
                 "gsc_compat_attach", "SXP_make_random_crosses", "SXP_self_n_times", "SXP_make_doubled_haploids", "SXP_make_clones",
                 "SXP_break_group_by_GEBV_num", "SXP_break_group_by_GEBV_percent", "SXP_create_markerblocks_nperchr",
                 "SXP_see_local_gebvs", "SXP_save_genotypes", "SXP_save_GEBVs", "SXP_save_local_GEBVs", "SXP_make_targeted_crosses",
                 "SXP_break_group_into_families", "SXP_break_group_into_halfsib_families", "SXP_break_group_evenly"):
        getattr(L, name).restype = VP
    L.shim_int_vector.argtypes = [C.c_ssize_t, VP]
    L.shim_real_vector.argtypes = [C.c_ssize_t, VP]
    L.shim_string.argtypes = [C.c_char_p]
    L.shim_extptr.argtypes = [VP]
    for name in ("shim_len", "shim_nrow", "shim_ncol", "shim_type"):
        getattr(L, name).argtypes = [VP]
    L.shim_len.restype = C.c_ssize_t
    L.shim_data.argtypes = [VP]
    return L


class R:
    """SEXP arguments the way R would pass them"""

    def __init__(self, L):
        self.L = L

    def ints(self, *v):
        a = np.asarray(v, dtype=np.int32)
        return VP(self.L.shim_int_vector(len(a), a.ctypes.data))

    def real(self, v):
        a = np.asarray([v], dtype=np.float64)
        return VP(self.L.shim_real_vector(1, a.ctypes.data))

    def lgl(self, v):
        return VP(self.L.shim_logical(int(v)))

    def chr(self, s):
        return VP(self.L.shim_string(s.encode() if s is not None else None))

    @property
    def null(self):
        return VP(self.L.shim_nil())

    def int_result(self, s):
        n = self.L.shim_len(s)
        return np.ctypeslib.as_array(C.cast(self.L.shim_data(s), C.POINTER(C.c_int)), shape=(n,)).copy()

    def genoptions(self, name=False, prefix=None, family=1, pedigree=True, ids=True, file_prefix=None, save_ped=False, save_bv=0,
                   save_geno=False, retain=True):
        """the ten trailing arguments of every crossing wrapper (SXP_create_genoptions, r-wrappers.c:1863-1912)"""
        return [self.lgl(name), self.chr(prefix) if prefix is not None else self.null, self.ints(family), self.lgl(pedigree), self.lgl(ids),
                self.chr(file_prefix) if file_prefix is not None else self.null, self.lgl(save_ped), self.ints(save_bv), self.lgl(save_geno),
                self.lgl(retain)]


def test_unmodified_r_wrappers_drive_the_engine(W, tmp_path):
    from genomicsimulation_b200 import SimData
    ds = datasets.synthetic(20, 900, 3, 100.0, seed=71, spacing="random", n_effect_sets=2)
    sim = SimData.from_dataset(ds)
    sim.use_native_draws(5)
    r = R(W)
    facade = VP(W.gsc_compat_attach(sim.d))
    exd = VP(W.shim_extptr(facade))

    # make.random.crosses(group 1, n.crosses = 30, offspring = 2, give.names with prefix "k"), map 0 = "the first map"
    out = VP(W.SXP_make_random_crosses(exd, r.ints(1), r.ints(30), r.ints(0), r.ints(0), *r.genoptions(name=True, prefix="k", family=2)))
    f1 = int(r.int_result(out)[0])
    assert f1 == 2 and sim.group_size(f1) == 60
    x = sim.export_group(f1, genotypes=False)
    assert x["ids"].tolist() == list(range(21, 81)) and sim.group_names(f1)[:2] == ["k21", "k22"]
    assert set(x["ped1"]) <= set(range(1, 21)) and (x["ped1"] != x["ped2"]).all()

    # self.n.times(f1, n = 3) and make.doubled.haploids on a VECTOR of groups (the wrappers loop, r-wrappers.c:2171-2186)
    s3 = int(r.int_result(VP(W.SXP_self_n_times(exd, r.ints(f1), r.ints(3), r.ints(0), *r.genoptions())))[0])
    dh = r.int_result(VP(W.SXP_make_doubled_haploids(exd, r.ints(f1, s3), r.ints(0), *r.genoptions())))
    assert sim.group_size(s3) == 60 and [sim.group_size(int(g)) for g in dh] == [60, 60]
    g = sim.export_group(int(dh[0]))["genotypes"]
    assert (g[:, 0::2] == g[:, 1::2]).all()                                   # doubled haploids are homozygous

    # select.by.gebv(f1, number = 7) == the seven best GEBVs; percent = 10 -> floor(60 * 10 / 100) = 6
    bv = sim.calculate_bvs(f1, 1)
    idx = sim.group_indexes(f1)
    top = int(r.int_result(VP(W.SXP_break_group_by_GEBV_num(exd, r.ints(f1), r.ints(0), r.ints(7), r.lgl(False))))[0])
    assert sorted(sim.group_indexes(top).tolist()) == sorted(idx[np.argsort(-bv, kind="stable")[:7]].tolist())
    low = int(r.int_result(VP(W.SXP_break_group_by_GEBV_percent(exd, r.ints(f1), r.ints(2), r.real(10.0), r.lgl(True))))[0])
    assert sim.group_size(low) == 5 or sim.group_size(low) == 6               # 53 left in f1 after the first selection: floor(5.3)

    # create.markerblocks + see.local.GEBVs: R's column-major REALSXP matrix from gsc_DecimalMatrix's row pointers
    exb = VP(W.SXP_create_markerblocks_nperchr(exd, r.ints(4), r.ints(0)))
    m = VP(W.SXP_see_local_gebvs(exd, exb, r.ints(s3), r.ints(2)))
    nrow, ncol = W.shim_nrow(m), W.shim_ncol(m)
    assert (nrow, ncol) == (120, 12)
    got = np.ctypeslib.as_array(C.cast(W.shim_data(m), C.POINTER(C.c_double)), shape=(ncol, nrow)).T
    want = sim.calculate_local_bvs(s3, sim.blocks_evenlength(4, 1), 2)
    assert np.array_equal(got, want)

    # the savers, through the wrappers: same bytes as the mirror's own savers
    W.SXP_save_genotypes(exd, r.chr(str(tmp_path / "w-geno.txt")), r.ints(top), r.lgl(True))
    sim.save_genotypes(tmp_path / "m-geno.txt", top, True)
    W.SXP_save_GEBVs(exd, r.chr(str(tmp_path / "w-bv.txt")), r.ints(f1), r.ints(1))
    sim.save_bvs(tmp_path / "m-bv.txt", f1, 1)
    W.SXP_save_local_GEBVs(exd, r.chr(str(tmp_path / "w-local.txt")), exb, r.ints(top), r.ints(1))
    sim.save_local_bvs(tmp_path / "m-local.txt", top, sim.blocks_evenlength(4, 1), 1, True)
    for n in ("geno", "bv", "local"):
        assert (tmp_path / ("w-%s.txt" % n)).read_bytes() == (tmp_path / ("m-%s.txt" % n)).read_bytes() and (tmp_path / ("w-%s.txt" % n)).stat().st_size > 100

    # make.targeted.crosses by NAME: gsc_get_index_of_name(d->m, ...) through the facade's handle (r-wrappers.c:2009-2013)
    n1 = W.shim_string(b"k21")
    tc = int(r.int_result(VP(W.SXP_make_targeted_crosses(exd, VP(n1), r.chr("k30"), r.ints(0), r.ints(0), *r.genoptions(family=3))))[0])
    y = sim.export_group(tc, genotypes=False)
    assert sim.group_size(tc) == 3 and set(y["ped1"]) == {21} and set(y["ped2"]) == {30}

    # break.group.into.families / .halfsib.families / .evenly: the group-split utilities behind their unmodified wrappers
    # (r-wrappers.c:1383-1517): what is left of f1 are full-sib pairs -> one group per pair of parents still together
    left = sim.export_group(f1, genotypes=False)
    pairs = {(a, b) for a, b in zip(left["ped1"].tolist(), left["ped2"].tolist())}
    fams = r.int_result(VP(W.SXP_break_group_into_families(exd, r.ints(f1))))
    assert len(fams) == len(pairs) and sim.group_size(f1) == 0
    assert sum(sim.group_size(int(g)) for g in fams) == len(left["ids"])
    for g in fams[:5]:
        z = sim.export_group(int(g), genotypes=False)
        assert len(set(zip(z["ped1"].tolist(), z["ped2"].tolist()))) == 1
    firsts = set(sim.export_group(s3, genotypes=False)["ped1"].tolist())
    hs = r.int_result(VP(W.SXP_break_group_into_halfsib_families(exd, r.ints(s3), r.ints(1))))
    assert len(hs) == len(firsts)
    assert all(len(set(sim.export_group(int(g), genotypes=False)["ped1"].tolist())) == 1 for g in hs)
    ev = r.int_result(VP(W.SXP_break_group_evenly(exd, r.ints(int(dh[0])), r.ints(4))))
    assert sorted(sim.group_size(int(g)) for g in ev) == [15, 15, 15, 15]
    sim.close()
