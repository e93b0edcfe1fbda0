"""CPU-side checks of the boundary: the C-ABI libraries load and export every symbol the headers
declare, and the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, txt)))


def test_engine_exports_every_declared_symbol():
    import genomicsimulation_b200 as g
    lib = C.CDLL(g.LIB_ENGINE)
    names = declared("gsb200.h", "gsb200_")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libgsb200.so does not export %s" % n
    assert set(names) == set(g._lib.ENGINE_SYMBOLS), "ctypes table out of step with include/gsb200.h"
    assert g.engine().gsb200_abi_version() == 1


def test_host_mirror_exports_every_declared_symbol():
    import genomicsimulation_b200 as g
    g.engine()
    lib = C.CDLL(g.LIB_HOST)
    names = declared("gsb200_gsc.h", "gsb_")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libgsb200_gsc.so does not export %s" % n
    assert set(names) == set(g._lib.HOST_SYMBOLS)
    opt = g.GenOptions.in_dll(lib, "GSB_BASIC_OPT")          # GSC_BASIC_OPT, sim-operations.c:11-23
    assert opt.family_size == 1 and opt.will_allocate_ids and opt.will_save_to_simdata and not opt.will_track_pedigree


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import genomicsimulation_b200 as g
    E = g.engine()
    ctx = C.c_void_p()
    st = E.gsb200_create(C.byref(ctx), 0, 100)
    assert st == 1 and not ctx.value                          # GSB200_ERR_CUDA
    assert b"no CPU fallback" in E.gsb200_last_error()
    with pytest.raises(g.EngineError):
        g.SimData(100)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "genomicsimulation_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "gs_oracle" not in txt and "libgsoracle" not in txt, f


def _compile_capi(out):
    import subprocess
    lib = os.path.join(ROOT, "genomicsimulation_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "test_capi.c"), "-L" + lib, "-lgsb200_gsc", "-lgsb200",
                           "-Wl,-rpath," + lib, "-lm", "-o", out])


def test_headers_are_plain_c_and_the_c_client_links(tmp_path):
    """include/*.h compile as strict C11 and a pure-C client of the mirror API links against the two
    shared objects (what r-wrappers.c would do)."""
    import genomicsimulation_b200 as g
    g.host()
    _compile_capi(str(tmp_path / "test_capi"))


@pytest.mark.gpu
def test_c_client_runs_a_breeding_scheme(tmp_path):
    import subprocess
    exe = str(tmp_path / "test_capi")
    _compile_capi(exe)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "capi ok: 216 genotypes" in out.stdout


def test_unmodified_r_wrappers_link_against_the_compat_header():
    """oracle/_ref/libgswrap.so (oracle/Makefile `wrappers`): hot-path functions of the reference's own
    src/r-wrappers.c, compiled unmodified against include/gsb200_compat.h.  Loading it here proves every
    gsc_* symbol they call is exported by libgsb200_gsc.so (no compute without a GPU)."""
    import ctypes as C
    import pytest
    path = os.path.join(ROOT, "oracle", "_ref", "libgswrap.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libgswrap.so not built here")
    from genomicsimulation_b200 import _lib
    _lib.host()
    L = C.CDLL(path)
    for name in ("SXP_make_random_crosses", "SXP_make_random_crosses_between", "SXP_make_targeted_crosses",
                 "SXP_make_all_unidirectional_crosses", "SXP_self_n_times", "SXP_make_doubled_haploids", "SXP_make_clones",
                 "SXP_break_group_by_GEBV_num", "SXP_break_group_by_GEBV_percent", "SXP_see_local_gebvs",
                 "SXP_create_markerblocks_nperchr", "SXP_save_genotypes", "SXP_save_allele_counts", "SXP_save_pedigrees",
                 "SXP_save_GEBVs", "SXP_save_local_GEBVs", "gsc_compat_attach", "gsc_make_random_crosses", "gsc_calculate_local_bvs"):
        assert getattr(L, name) is not None
