"""The host mirror's bookkeeping (storage order, ids, names, groups, the free-row list) model-checked on CPU.

`genomicsimulation_b200/host/gsb_simdata.c` is compiled against `tests/c/stub_engine.c` — a stand-in for the CUDA engine in
which every device call succeeds at once and does nothing — so that what the mirror keeps on the HOST can be driven through
long random sequences of the reference's operations without a GPU and compared, after every step, with a plain Python model of
what gsc_SimData would hold (core.c:2810-2841 make_group_from, 3937-3957 new group numbers, 4658-4692 delete_group,
8168-8180 / 8307-8335 ids and names of a batch).  The genotypes themselves are the GPU tests' business (tests/test_gpu_*.py).
"""
import ctypes as C
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = 32


class GenOptions(C.Structure):      # gsb_GenOptions, include/gsb200_gsc.h
    _fields_ = [("will_name_offspring", C.c_bool), ("offspring_name_prefix", C.c_char_p), ("family_size", C.c_uint),
                ("will_track_pedigree", C.c_bool), ("will_allocate_ids", C.c_bool), ("filename_prefix", C.c_char_p),
                ("will_save_pedigree_to_file", C.c_bool), ("will_save_bvs_to_file", C.c_uint),
                ("will_save_alleles_to_file", C.c_bool), ("will_save_to_simdata", C.c_bool)]


def opts(**kw):
    g = GenOptions(False, None, 1, False, True, None, False, 0, False, True)
    for k, v in kw.items():
        setattr(g, k, v)
    return g


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("host_stub")
    cc = os.environ.get("CC", "gcc")
    extra = os.environ.get("GSB_TEST_CFLAGS", "").split()      # e.g. -fsanitize=address with LD_PRELOAD=libasan.so
    # mirror and stub in ONE shared object, bound to each other (-Bsymbolic): the real libgsb200.so may be loaded in this process too
    subprocess.check_call([cc, "-O2", "-g", "-std=gnu11", "-fPIC", "-shared", "-w", *extra, "-I" + os.path.join(ROOT, "include"),
                           "-o", str(out / "libhost.so"), os.path.join(ROOT, "genomicsimulation_b200/host/gsb_simdata.c"),
                           os.path.join(ROOT, "tests/c/stub_engine.c"), "-lm", "-Wl,-Bsymbolic"])
    L = C.CDLL(str(out / "libhost.so"))
    vp, u = C.c_void_p, C.c_uint
    L.gsb_create_simdata.restype = vp
    L.gsb_create_simdata.argtypes = [C.c_int, u]
    L.gsb_delete_simdata.argtypes = [vp]
    L.gsb_set_print.argtypes = [vp, vp]
    L.gsb_load_genotypes.restype = u
    L.gsb_load_genotypes.argtypes = [vp, u, C.c_char_p, C.POINTER(C.c_char_p)]
    L.gsb_load_map.restype = u
    L.gsb_load_map.argtypes = [vp, u, vp, vp, vp, vp, vp]
    L.gsb_make_clones.restype = u
    L.gsb_make_clones.argtypes = [vp, u, C.c_bool, GenOptions]
    L.gsb_make_random_crosses.restype = u
    L.gsb_make_random_crosses.argtypes = [vp, u, u, u, u, GenOptions]
    L.gsb_delete_group.argtypes = [vp, u]
    L.gsb_make_group_from.restype = u
    L.gsb_make_group_from.argtypes = [vp, u, vp]
    L.gsb_combine_groups.restype = u
    L.gsb_combine_groups.argtypes = [vp, u, vp]
    L.gsb_get_n_genotypes.restype = u
    L.gsb_get_n_genotypes.argtypes = [vp]
    L.gsb_get_current_id.restype = u
    L.gsb_get_current_id.argtypes = [vp]
    L.gsb_get_group_data.restype = u
    L.gsb_get_group_data.argtypes = [vp, u, u, vp, vp, vp, vp, vp, vp]
    L.gsb_get_group_member_name.restype = C.c_char_p
    L.gsb_get_group_member_name.argtypes = [vp, u, u]
    return L


class Model:
    """what the reference would hold: one record per genotype in storage order"""

    def __init__(self):
        self.g = []          # [id, name or None, group]
        self.current_id = 0

    def new_group(self):
        used = {r[2] for r in self.g}
        k = 1
        while k in used:
            k += 1
        return k

    def append(self, n, group, names=None, name_prefix=None, ids=True):
        for i in range(n):
            self.current_id += 1 if ids else 0
            ident = self.current_id if ids else 0
            name = names[i] if names else (("%s%u" % (name_prefix, ident)) if name_prefix is not None else None)
            self.g.append([ident, name, group])


def snapshot(L, d):
    n = L.gsb_get_n_genotypes(d)
    ids, groups = (C.c_uint * max(n, 1))(), (C.c_uint * max(n, 1))()
    assert L.gsb_get_group_data(d, 0, n, None, ids, None, None, groups, None) == n
    names = [L.gsb_get_group_member_name(d, 0, i) for i in range(n)]
    return [[ids[i], names[i].decode() if names[i] else None, groups[i]] for i in range(n)]


@pytest.mark.parametrize("seed,native", [(1, False), (2, False), (3, False), (4, True), (5, True)])
def test_random_operation_sequences_match_the_model(lib, seed, native):
    """native: GSB_DRAWS_NATIVE, where random crosses take the batched device path (scaffold_device_pairs: the picks come back
    from the engine — here from the stub — and the pedigrees are filled from them)"""
    L = lib
    rng = random.Random(seed)
    d = L.gsb_create_simdata(0, M)
    assert d
    L.gsb_set_print(d, None)
    if native:
        L.gsb_set_draw_mode.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.gsb_set_draw_mode(d, 1, 99)
    lam = (C.c_double * 1)(1.0)
    off = (C.c_uint * 2)(0, M)
    dists = (C.c_double * M)(*[i / M for i in range(M)])
    idx = (C.c_uint * M)(*range(M))
    has = (C.c_ubyte * M)(*([1] * M))
    assert L.gsb_load_map(d, 1, lam, off, dists, idx, has) == 1
    m = Model()
    chars = b"A" * (2 * M * 4000)
    serial = 0
    for step in range(250):
        groups = sorted({r[2] for r in m.g})
        op = rng.choice(["load", "load_named", "clone", "cross", "cross_named", "delete", "delete", "group_from", "combine", "cross_drop"])
        if op in ("load", "load_named") or not groups:
            n = rng.choice([1, 3, 7, 1500])
            names = None
            if op == "load_named":
                names = ["L%d_%d" % (serial, i) if rng.random() < 0.8 else None for i in range(n)]
                serial += 1
            arr = (C.c_char_p * n)(*[s.encode() if s else None for s in names]) if names else None
            g = L.gsb_load_genotypes(d, n, chars, arr)
            assert g == m.new_group()
            m.append(n, g, names=names if names else None, ids=True) if not names else [m.append(1, g, names=[s]) for s in names]
        elif op == "clone":
            src = rng.choice(groups)
            members = [r for r in m.g if r[2] == src]
            inherit = rng.random() < 0.7
            g = L.gsb_make_clones(d, src, inherit, opts())
            assert g == m.new_group()
            for r in members:
                m.append(1, g, names=[r[1] if inherit else None])
        elif op in ("cross", "cross_named", "cross_drop"):
            src = rng.choice(groups)
            size = sum(1 for r in m.g if r[2] == src)
            n = rng.choice([1, 5, 1200, 2300])
            o = opts(will_name_offspring=(op == "cross_named"), offspring_name_prefix=b"X" if op == "cross_named" else None,
                     will_save_to_simdata=(op != "cross_drop"), will_track_pedigree=native)
            g = L.gsb_make_random_crosses(d, src, n, 0, 0, o)
            if size < 2:
                assert g == 0        # the reference refuses to cross a group with itself when it has one member
            elif op == "cross_drop":
                assert g == 0
                m.current_id = L.gsb_get_current_id(d)       # ids of offspring that were not kept are spent all the same
            else:
                assert g == m.new_group()
                m.append(n, g, name_prefix="X" if op == "cross_named" else None)
        elif op == "delete":
            g = rng.choice(groups)
            L.gsb_delete_group(d, g)
            m.g = [r for r in m.g if r[2] != g]
        elif op == "group_from":
            n = len(m.g)
            k = rng.randint(1, min(n, 40))
            picks = sorted(rng.sample(range(n), k)) if rng.random() < 0.5 else list(range(rng.randint(0, n - k), n))[:k]
            arr = (C.c_uint * len(picks))(*picks)
            want = m.new_group()
            g = L.gsb_make_group_from(d, len(picks), arr)
            assert g == want
            for i in picks:
                m.g[i][2] = g
        elif op == "combine" and len(groups) >= 2:
            a, b = rng.sample(groups, 2)
            arr = (C.c_uint * 2)(a, b)
            assert L.gsb_combine_groups(d, 2, arr) == a
            for r in m.g:
                if r[2] == b:
                    r[2] = a
        got = snapshot(L, d)
        assert len(got) == len(m.g), (step, op)
        assert got == m.g, (step, op, next((i, x, y) for i, (x, y) in enumerate(zip(got, m.g)) if x != y))
        if len(m.g) > 6000:          # keep the sequences fast: drop the largest group
            big = max(groups, key=lambda q: sum(1 for r in m.g if r[2] == q))
            L.gsb_delete_group(d, big)
            m.g = [r for r in m.g if r[2] != big]
    L.gsb_delete_simdata(d)


def test_sharded_generations_keep_the_books_of_one_process(lib):
    """gsb_shard_select_and_cross + gsb_delete_group for a few generations (one rank): ids, names, groups and — from the
    stub's known picks — pedigrees; generations that are one run of the storage order and generations that are not."""
    L = lib
    vp, u, ull = C.c_void_p, C.c_uint, C.c_ulonglong
    L.gsb_set_draw_mode.argtypes = [vp, C.c_int, C.c_uint64]
    L.gsb_load_effects.restype = u
    L.gsb_load_effects.argtypes = [vp, vp, C.c_char_p, vp, vp]
    L.gsb_shard_init.argtypes = [vp, u, u, ull, u]
    L.gsb_shard_handle.argtypes = [vp, vp]
    L.gsb_shard_connect.argtypes = [vp, vp]
    L.gsb_shard_select_and_cross.restype = u
    L.gsb_shard_select_and_cross.argtypes = [vp, u, ull, u, u, C.c_bool, ull, u, GenOptions]
    L.gsb_shard_get_local_bvs.restype = u
    L.gsb_shard_get_local_bvs.argtypes = [vp, u, vp]
    n, top = 3000, 30
    d = L.gsb_create_simdata(0, M)
    L.gsb_set_print(d, None)
    L.gsb_set_draw_mode(d, 1, 5)
    lam = (C.c_double * 1)(1.0)
    off = (C.c_uint * 2)(0, M)
    dists = (C.c_double * M)(*[i / M for i in range(M)])
    idx = (C.c_uint * M)(*range(M))
    has = (C.c_ubyte * M)(*([1] * M))
    assert L.gsb_load_map(d, 1, lam, off, dists, idx, has) == 1
    cum = (C.c_uint * (M + 1))(*range(M + 1))
    eff = (C.c_double * M)(*([0.1] * M))
    e = L.gsb_load_effects(d, cum, b"A" * M, eff, None)
    assert e == 1
    chars = b"A" * (2 * M * n)
    m = Model()
    assert L.gsb_load_genotypes(d, 7, chars, (C.c_char_p * 7)(*[b"f%d" % i for i in range(7)])) == 1      # named founders in front
    for i in range(7):
        m.append(1, 1, names=["f%d" % i])
    g = L.gsb_load_genotypes(d, n, chars, None)
    m.append(n, g)
    handle = C.create_string_buffer(128)
    assert L.gsb_shard_init(d, 0, 1, n, top) == 0 and L.gsb_shard_handle(d, handle) == 0 and L.gsb_shard_connect(d, handle) == 0
    bv = (C.c_double * n)()
    k1 = [(i * 2654435761) % top for i in range(n)]       # tests/c/stub_engine.c: 64-bit products
    k2 = [(i * 40503 + 7) % top for i in range(n)]
    for gen in range(7):
        named = gen in (2, 3)
        if gen == 5:            # more than 63 groups in use: the new group's number no longer fits the one-pass bit mask
            for _ in range(70):
                q = L.gsb_load_genotypes(d, 1, chars, None)
                assert q == m.new_group()
                m.append(1, q)
        if gen == 4:            # break the run: the generation's members are no longer contiguous in storage order
            extra = L.gsb_load_genotypes(d, 2, chars, None)
            m.append(2, extra)
            arr = (C.c_uint * 1)(len(m.g) - 1)
            moved = L.gsb_make_group_from(d, 1, arr)
            first = next(i for i, r in enumerate(m.g) if r[2] == g)
            L.gsb_delete_group(d, moved)
            m.g.pop()
            # swap one member out and another genotype in, keeping the size
            arr = (C.c_uint * 1)(first)
            gone = L.gsb_make_group_from(d, 1, arr)
            m.g[first][2] = gone
            arr = (C.c_uint * 2)(next(i for i, r in enumerate(m.g) if r[2] == g), len(m.g) - 1)
            g2 = L.gsb_make_group_from(d, 2, arr)
            for i in (arr[0], arr[1]):
                m.g[i][2] = g2
            arr = (C.c_uint * 2)(g, g2)
            assert L.gsb_combine_groups(d, 2, arr) == g
            for r in m.g:
                if r[2] == g2:
                    r[2] = g
            assert sum(1 for r in m.g if r[2] == g) == n
        o = opts(will_track_pedigree=True, will_name_offspring=named, offspring_name_prefix=b"G" if named else None)
        want = m.new_group()
        new = L.gsb_shard_select_and_cross(d, g, n, e, top, False, n, 0, o)
        assert new == want, gen
        assert L.gsb_shard_get_local_bvs(d, n, bv) == n
        m.append(n, new, name_prefix="G" if named else None)
        got = snapshot(L, d)
        assert got == m.g, (gen, next((i, x, y) for i, (x, y) in enumerate(zip(got, m.g)) if x != y))
        p1, p2 = (C.c_uint * n)(), (C.c_uint * n)()
        assert L.gsb_get_group_data(d, new, n, None, None, p1, p2, None, None) == n
        assert list(p1) == [k + 1 for k in k1] and list(p2) == [k + 1 for k in k2], gen
        L.gsb_delete_group(d, g)
        m.g = [r for r in m.g if r[2] != g]
        g = new
        assert snapshot(L, d) == m.g, gen
    L.gsb_delete_simdata(d)
