"""oracle/port.py — TEST INFRASTRUCTURE ONLY.

ctypes driver for oracle/libgsoracle.so, the plain-C restatement of the
reference hot path (oracle/gs_oracle.c).  Method names mirror oracle/ref.py so
tests can run the same scenario through either oracle.  Map/effect arguments are
INDEXES here (the restatement has no id→index table; ids are `index + 1`).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsoracle.so")

U = C.c_uint
DRAW_U, DRAW_P, DRAW_G, DRAW_E = 1, 2, 3, 4


class GenOptions(C.Structure):
    _fields_ = [("family_size", C.c_uint), ("track_pedigree", C.c_int), ("allocate_ids", C.c_int),
                ("name_offspring", C.c_int), ("name_prefix", C.c_char_p), ("retain", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.gso_new.restype = C.c_void_p
        L.gso_member_name.restype = C.c_char_p
        L.gsdraw_tape_len.restype = C.c_size_t
        L.gsdraw_replay_stop.restype = C.c_size_t
        L.gsdraw_seed.argtypes = [C.c_uint64]
        for nm in ("gso_make_random_crosses", "gso_make_random_crosses_between", "gso_make_targeted_crosses",
                   "gso_make_all_unidirectional_crosses", "gso_self_n_times", "gso_make_doubled_haploids",
                   "gso_make_clones"):
            getattr(L, nm).restype = C.c_uint
        L.gso_make_random_crosses.argtypes = [C.c_void_p, U, U, U, U, GenOptions]
        L.gso_make_random_crosses_between.argtypes = [C.c_void_p, U, U, U, U, U, U, U, GenOptions]
        L.gso_make_targeted_crosses.argtypes = [C.c_void_p, U, C.c_void_p, C.c_void_p, U, U, GenOptions]
        L.gso_make_all_unidirectional_crosses.argtypes = [C.c_void_p, U, U, GenOptions]
        L.gso_self_n_times.argtypes = [C.c_void_p, U, U, U, GenOptions]
        L.gso_make_doubled_haploids.argtypes = [C.c_void_p, U, U, GenOptions]
        L.gso_make_clones.argtypes = [C.c_void_p, U, C.c_int, GenOptions]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _opts(family_size=1, track_pedigree=True, allocate_ids=True, name_offspring=False,
          retain=True, name_prefix=None):
    return GenOptions(family_size, int(track_pedigree), int(allocate_ids), int(name_offspring),
                      name_prefix.encode() if name_prefix is not None else None, int(retain))


def seed(s):
    lib().gsdraw_seed(C.c_uint64(s))


def record_start():
    lib().gsdraw_record(1)


def record_stop():
    from .ref import Tape
    L = lib()
    n = L.gsdraw_tape_len()
    vals = np.empty(n, dtype=np.float64)
    kinds = np.empty(n, dtype=np.uint8)
    L.gsdraw_tape_copy(_p(vals), _p(kinds))
    L.gsdraw_record(0)
    return Tape(vals, kinds)


class Replay:
    def __init__(self, tape):
        self.tape = tape
        self.consumed = None
        self.error = None

    def __enter__(self):
        lib().gsdraw_replay_start(_p(self.tape.values), _p(self.tape.kinds), C.c_size_t(len(self.tape)))
        return self

    def __exit__(self, *exc):
        L = lib()
        self.error = bool(L.gsdraw_replay_error())
        self.consumed = L.gsdraw_replay_stop()
        return False


class PortSim:
    def __init__(self, n_markers):
        self.L = lib()
        self.M = n_markers
        self.d = C.c_void_p(self.L.gso_new(n_markers))

    @classmethod
    def from_dataset(cls, ds):
        s = cls(ds["n_markers"])
        s.founders = s.add_genotypes(ds["genotypes"], ds.get("names"))
        for m in ds["maps"]:
            s.add_map(m)
        for e in ds["effects"]:
            s.add_effects(e)
        return s

    def close(self):
        if self.d:
            self.L.gso_free(self.d)
            self.d = None

    # ---- setup -----------------------------------------------------------
    def add_genotypes(self, g, names=None):
        g = np.ascontiguousarray(g, dtype=np.uint8)
        n = g.shape[0]
        arr = None
        if names is not None:
            arr = (C.c_char_p * n)(*[nm.encode() if nm is not None else None for nm in names])
        return self.L.gso_add_genotypes(self.d, n, _p(g), arr)

    def add_map(self, m):
        lam = np.ascontiguousarray(m["lam"], dtype=np.float64)
        off = np.ascontiguousarray(m["chr_offset"], dtype=np.uint32)
        dists = np.ascontiguousarray(m["dists"], dtype=np.float64)
        mix = np.ascontiguousarray(m["marker_index"], dtype=np.uint32)
        has = np.ascontiguousarray(m["has_dists"], dtype=np.uint8)
        return self.L.gso_add_map(self.d, len(lam), _p(lam), _p(off), _p(dists), _p(mix), _p(has))

    def add_effects(self, e):
        cumn = np.ascontiguousarray(e["cumn"], dtype=np.uint32)
        allele = np.ascontiguousarray(e["allele"], dtype=np.uint8)
        eff = np.ascontiguousarray(e["eff"], dtype=np.float64)
        centre = None if e.get("centre") is None else np.ascontiguousarray(e["centre"], dtype=np.float64)
        return self.L.gso_add_effects(self.d, _p(cumn), _p(allele), _p(eff), _p(centre) if centre is not None else None)

    # scenario helpers (see oracle/ref.py): the restatement takes indexes directly
    def map_arg(self, map_index):
        return map_index

    def eff_arg(self, eff_index):
        return eff_index

    # ---- inspection --------------------------------------------------------
    @property
    def n_markers(self):
        return self.M

    @property
    def n_genotypes(self):
        return self.L.gso_n_genotypes(self.d)

    @property
    def current_id(self):
        return self.L.gso_current_id(self.d)

    def group_size(self, group):
        return self.L.gso_group_size(self.d, group)

    def export_group(self, group=0, genotypes=True):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        g = np.zeros((n, 2 * self.M), dtype=np.uint8) if genotypes else None
        ids, p1, p2, gr, gi = (np.zeros(n, dtype=np.uint32) for _ in range(5))
        got = self.L.gso_export_group(self.d, group, n, _p(g) if genotypes else None,
                                      _p(ids), _p(p1), _p(p2), _p(gr), _p(gi))
        assert got == n
        return dict(genotypes=g, ids=ids, ped1=p1, ped2=p2, groups=gr, indexes=gi)

    def group_indexes(self, group):
        return self.export_group(group, genotypes=False)["indexes"]

    def group_names(self, group):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = []
        for i in range(n):
            nm = self.L.gso_member_name(self.d, group, i)
            out.append(nm.decode() if nm is not None else None)
        return out

    # ---- crossing ----------------------------------------------------------
    def make_random_crosses(self, group, n, cap=0, map_index=0, **kw):
        return self.L.gso_make_random_crosses(self.d, group, n, cap, map_index, _opts(**kw))

    def make_random_crosses_between(self, g1, g2, n, cap1=0, cap2=0, map1=0, map2=0, **kw):
        return self.L.gso_make_random_crosses_between(self.d, g1, g2, n, cap1, cap2, map1, map2, _opts(**kw))

    def make_targeted_crosses(self, p1, p2, map1=0, map2=0, **kw):
        a = np.ascontiguousarray(p1, dtype=np.uint32)
        b = np.ascontiguousarray(p2, dtype=np.uint32)
        return self.L.gso_make_targeted_crosses(self.d, len(a), _p(a), _p(b), map1, map2, _opts(**kw))

    def make_all_unidirectional_crosses(self, group, map_index=0, **kw):
        return self.L.gso_make_all_unidirectional_crosses(self.d, group, map_index, _opts(**kw))

    def self_n_times(self, n, group, map_index=0, **kw):
        return self.L.gso_self_n_times(self.d, n, group, map_index, _opts(**kw))

    def make_doubled_haploids(self, group, map_index=0, **kw):
        return self.L.gso_make_doubled_haploids(self.d, group, map_index, _opts(**kw))

    def make_clones(self, group, inherit_names=False, **kw):
        return self.L.gso_make_clones(self.d, group, int(inherit_names), _opts(**kw))

    def generate_gamete(self, parent, map_index=0):
        p = np.ascontiguousarray(parent, dtype=np.uint8)
        out = np.zeros(2 * self.M, dtype=np.uint8)
        self.L.gso_generate_gamete(self.d, _p(p), _p(out), map_index)
        return out[0::2].copy()

    # ---- evaluation ----------------------------------------------------------
    def calculate_bvs(self, group, eff_index):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = np.zeros(n, dtype=np.float64)
        got = self.L.gso_calculate_bvs(self.d, group, eff_index, _p(out))
        return out[:got]

    def calculate_allele_counts(self, group, allele):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = np.zeros((n, self.M), dtype=np.float64)
        a = allele if isinstance(allele, bytes) else allele.encode()
        self.L.gso_calculate_allele_counts(self.d, group, C.c_char(a), _p(out))
        return out

    def calculate_local_bvs(self, group, offsets, markers, eff_index):
        o = np.ascontiguousarray(offsets, dtype=np.uint32)
        m = np.ascontiguousarray(markers, dtype=np.uint32)
        nb = len(o) - 1
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = np.zeros((2 * n, nb), dtype=np.float64)
        self.L.gso_calculate_local_bvs(self.d, group, nb, _p(o), _p(m), eff_index, _p(out))
        return out

    def blocks_evenlength(self, n, map_index=0, n_chr=None):
        if n_chr is None:
            raise ValueError("n_chr required")
        off = np.zeros(n * n_chr + 1, dtype=np.uint32)
        mk = np.zeros(max(self.M, 1), dtype=np.uint32)
        nb = self.L.gso_create_evenlength_blocks_each_chr(self.d, map_index, n, _p(off), _p(mk))
        return off[:nb + 1], mk[:off[nb]]

    def split_by_bv(self, group, eff_index, top_n, low_is_best=False):
        return self.L.gso_split_by_bv(self.d, group, eff_index, top_n, int(low_is_best))

    def make_group_from(self, indexes):
        a = np.ascontiguousarray(indexes, dtype=np.uint32)
        return self.L.gso_make_group_from(self.d, len(a), _p(a))

    def delete_group(self, group):
        self.L.gso_delete_group(self.d, group)
