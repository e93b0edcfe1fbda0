"""oracle/ref.py — TEST INFRASTRUCTURE ONLY.

ctypes driver for oracle/_ref/libgsref.so: the UNMODIFIED reference core
(/root/reference/src/sim-operations.c) compiled against the R shim, behind the
flat facade in oracle/ref_bridge.c.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgsref.so")

U = C.c_uint
_u_p = C.POINTER(C.c_uint)
_d_p = C.POINTER(C.c_double)

# kinds on the draw tape (oracle/draws.h)
DRAW_U, DRAW_P, DRAW_G, DRAW_E = 1, 2, 3, 4


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libgsref.so not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        L.refb_new.restype = C.c_void_p
        L.refb_marker_name.restype = C.c_char_p
        L.refb_group_member_name.restype = C.c_char_p
        L.refb_map_chr_info.restype = C.c_double
        for nm in ("refb_blocks_evenlength", "refb_blocks_from_file", "refb_blocks_from_csr"):
            getattr(L, nm).restype = C.c_void_p
        L.gsdraw_tape_len.restype = C.c_size_t
        L.gsdraw_replay_stop.restype = C.c_size_t
        L.shim_captured_len.restype = C.c_size_t
        L.gsdraw_seed.argtypes = [C.c_uint64]
        _lib = L
    return _lib


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def gen_options(family_size=1, track_pedigree=True, allocate_ids=True, name_offspring=False,
                retain=True, save_pedigree=False, save_bvs=0, save_genotypes=False):
    return (U * 8)(family_size, int(track_pedigree), int(allocate_ids), int(name_offspring),
                   int(retain), int(save_pedigree), int(save_bvs), int(save_genotypes))


class Tape:
    """A recorded sequence of draws: parallel arrays `values` (f64) and `kinds` (u8)."""

    def __init__(self, values, kinds):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.kinds = np.ascontiguousarray(kinds, dtype=np.uint8)

    def __len__(self):
        return len(self.values)

    def draws_only(self):
        keep = (self.kinds == DRAW_U) | (self.kinds == DRAW_P)
        return Tape(self.values[keep], self.kinds[keep])


def seed(s):
    lib().gsdraw_seed(C.c_uint64(s))


def record_start():
    lib().gsdraw_record(1)


def record_stop():
    L = lib()
    n = L.gsdraw_tape_len()
    vals = np.empty(n, dtype=np.float64)
    kinds = np.empty(n, dtype=np.uint8)
    L.gsdraw_tape_copy(_ptr(vals, C.c_double), _ptr(kinds, C.c_uint8))
    L.gsdraw_record(0)
    return Tape(vals, kinds)


class Replay:
    """Context manager: the reference consumes draws from `tape` instead of its RNG."""

    def __init__(self, tape):
        self.tape = tape
        self.consumed = None
        self.error = None

    def __enter__(self):
        lib().gsdraw_replay_start(_ptr(self.tape.values, C.c_double), _ptr(self.tape.kinds, C.c_uint8),
                                  C.c_size_t(len(self.tape)))
        return self

    def __exit__(self, *exc):
        L = lib()
        self.error = bool(L.gsdraw_replay_error())
        self.consumed = L.gsdraw_replay_stop()
        return False


def set_print_mode(mode):
    """0 discard, 1 stdout, 2 capture."""
    lib().shim_set_print_mode(mode)


def captured_text():
    L = lib()
    n = L.shim_captured_len()
    buf = C.create_string_buffer(n + 1)
    L.shim_captured_copy(buf)
    return buf.raw[:n].decode("latin-1")


class RefSim:
    """One reference SimData (core.h `gsc_SimData`)."""

    def __init__(self):
        self.L = lib()
        self.d = C.c_void_p(self.L.refb_new())

    def close(self):
        if self.d:
            self.L.refb_free(self.d)
            self.d = None

    # ---- loading -------------------------------------------------------
    def load(self, geno, mapf=None, eff=None):
        ids = (U * 3)()
        enc = lambda s: s.encode() if s is not None else None
        self.L.refb_load(self.d, enc(geno), enc(mapf), enc(eff), ids)
        return ids[0], ids[1], ids[2]

    def load_effects(self, f):
        return self.L.refb_load_effects(self.d, f.encode())

    def load_map(self, f):
        return self.L.refb_load_map(self.d, f.encode())

    def load_genotypes(self, f):
        return self.L.refb_load_genotypes(self.d, f.encode())

    # ---- inspection ----------------------------------------------------
    @property
    def n_markers(self):
        return self.L.refb_n_markers(self.d)

    @property
    def n_genotypes(self):
        return self.L.refb_n_genotypes(self.d)

    @property
    def current_id(self):
        return self.L.refb_current_id(self.d)

    def marker_names(self):
        return [self.L.refb_marker_name(self.d, m).decode() for m in range(self.n_markers)]

    def group_size(self, group):
        return self.L.refb_group_size(self.d, group)

    def export_group(self, group=0, genotypes=True):
        """Returns dict(genotypes (n, 2M) uint8 | None, ids, ped1, ped2, groups) in storage order."""
        n = self.n_genotypes if group == 0 else self.group_size(group)
        M = self.n_markers
        g = np.zeros((n, 2 * M), dtype=np.uint8) if genotypes else None
        ids, p1, p2, gr = (np.zeros(n, dtype=np.uint32) for _ in range(4))
        got = self.L.refb_export_group(self.d, group, n,
                                       g.ctypes.data_as(C.c_char_p) if genotypes else None,
                                       _ptr(ids, U), _ptr(p1, U), _ptr(p2, U), _ptr(gr, U))
        assert got == n, (got, n)
        return dict(genotypes=g, ids=ids, ped1=p1, ped2=p2, groups=gr)

    def group_indexes(self, group):
        n = self.group_size(group)
        out = np.zeros(n, dtype=np.uint32)
        self.L.refb_group_indexes(self.d, group, n, _ptr(out, U))
        return out

    def group_names(self, group):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        res = []
        for i in range(n):
            nm = self.L.refb_group_member_name(self.d, group, i)
            res.append(nm.decode() if nm is not None else None)
        return res

    def export_map(self, map_index=0):
        """Returns list of dict(type, lam, n, first, dists (f64 or None), marker_indexes)."""
        chrs = []
        for c in range(self.L.refb_map_n_chr(self.d, map_index)):
            info = (U * 3)()
            lam = self.L.refb_map_chr_info(self.d, map_index, c, info)
            n = info[1]
            dists = np.full(n, np.nan, dtype=np.float64)
            mi = np.zeros(n, dtype=np.uint32)
            self.L.refb_map_chr_arrays(self.d, map_index, c, _ptr(dists, C.c_double), _ptr(mi, U))
            has = bool(self.L.refb_map_chr_has_dists(self.d, map_index, c))
            chrs.append(dict(type=info[0], lam=lam, n=n, first=info[2],
                             dists=dists if has else None, marker_indexes=mi))
        return chrs

    def map_id(self, map_index):
        return self.L.refb_map_id(self.d, map_index)

    # scenario helpers: the argument each implementation expects for "the i-th map / effect set"
    def map_arg(self, map_index):
        return self.map_id(map_index)

    def eff_arg(self, eff_index):
        return self.eff_id(eff_index)

    def export_effects(self, eff_index=0):
        M = self.n_markers
        nnz = self.L.refb_eff_nnz(self.d, eff_index)
        cumn = np.zeros(M, dtype=np.uint32)
        allele = np.zeros(nnz, dtype=np.uint8)
        eff = np.zeros(nnz, dtype=np.float64)
        centre = np.zeros(M, dtype=np.float64)
        has = self.L.refb_eff_export(self.d, eff_index, _ptr(cumn, U), allele.ctypes.data_as(C.c_char_p),
                                     _ptr(eff, C.c_double), _ptr(centre, C.c_double))
        return dict(cumn=cumn, allele=allele, eff=eff, centre=centre if has else None)

    def eff_id(self, eff_index):
        return self.L.refb_eff_id(self.d, eff_index)

    def set_centres(self, eff_id, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        return self.L.refb_set_centres(self.d, eff_id, len(v), _ptr(v, C.c_double))

    # ---- crossing ------------------------------------------------------
    @staticmethod
    def _pfx(kw):
        np_ = kw.pop("name_prefix", None)
        fp_ = kw.pop("file_prefix", None)
        return (np_.encode() if np_ is not None else None, fp_.encode() if fp_ is not None else None)

    def make_random_crosses(self, group, n, cap=0, map_id=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_random_crosses(self.d, group, n, cap, map_id, gen_options(**kw), np_, fp_)

    def make_random_crosses_between(self, g1, g2, n, cap1=0, cap2=0, map1=0, map2=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_random_crosses_between(self.d, g1, g2, n, cap1, cap2, map1, map2,
                                                       gen_options(**kw), np_, fp_)

    def make_targeted_crosses(self, p1, p2, map1=0, map2=0, **kw):
        np_, fp_ = self._pfx(kw)
        a, b = _u32(p1), _u32(p2)
        return self.L.refb_make_targeted_crosses(self.d, len(a), _ptr(a, U), _ptr(b, U), map1, map2,
                                                 gen_options(**kw), np_, fp_)

    def make_crosses_from_file(self, path, map1=0, map2=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_crosses_from_file(self.d, path.encode(), map1, map2, gen_options(**kw), np_, fp_)

    def make_double_crosses_from_file(self, path, map1=0, map2=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_double_crosses_from_file(self.d, path.encode(), map1, map2, gen_options(**kw), np_, fp_)

    def make_all_unidirectional_crosses(self, group, map_id=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_all_unidirectional_crosses(self.d, group, map_id, gen_options(**kw), np_, fp_)

    def self_n_times(self, n, group, map_id=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_self_n_times(self.d, n, group, map_id, gen_options(**kw), np_, fp_)

    def make_doubled_haploids(self, group, map_id=0, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_doubled_haploids(self.d, group, map_id, gen_options(**kw), np_, fp_)

    def make_clones(self, group, inherit_names=False, **kw):
        np_, fp_ = self._pfx(kw)
        return self.L.refb_make_clones(self.d, group, int(inherit_names), gen_options(**kw), np_, fp_)

    def generate_gamete(self, parent, map_index=0):
        """parent: (2M,) uint8. Returns (M,) uint8 gamete."""
        M = self.n_markers
        p = np.ascontiguousarray(parent, dtype=np.uint8)
        out = np.zeros(2 * M, dtype=np.uint8)
        self.L.refb_generate_gamete(self.d, p.ctypes.data_as(C.c_char_p), out.ctypes.data_as(C.c_char_p), map_index)
        return out[0::2].copy()

    # ---- evaluation ----------------------------------------------------
    def calculate_bvs(self, group, eff_id):
        n = self.L.refb_calculate_bvs(self.d, group, eff_id, 0, None)
        out = np.zeros(n, dtype=np.float64)
        self.L.refb_calculate_bvs(self.d, group, eff_id, n, _ptr(out, C.c_double))
        return out

    def blocks_evenlength(self, n, map_id=0):
        return RefBlocks(self, C.c_void_p(self.L.refb_blocks_evenlength(self.d, map_id, n)))

    def blocks_from_file(self, f):
        return RefBlocks(self, C.c_void_p(self.L.refb_blocks_from_file(self.d, f.encode())))

    def blocks_from_csr(self, offsets, markers):
        o, m = _u32(offsets), _u32(markers)
        return RefBlocks(self, C.c_void_p(self.L.refb_blocks_from_csr(len(o) - 1, _ptr(o, U), _ptr(m, U))))

    def calculate_local_bvs(self, group, blocks, eff_id):
        nb = blocks.n_blocks
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = np.zeros((2 * n, nb), dtype=np.float64)
        rows = self.L.refb_calculate_local_bvs(self.d, group, blocks.h, eff_id, 2 * n, _ptr(out, C.c_double))
        assert rows == 2 * n or nb == 0, (rows, n)
        return out

    def calculate_allele_counts(self, group, allele):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = np.zeros((n, self.n_markers), dtype=np.float64)
        a = allele if isinstance(allele, bytes) else allele.encode()
        rows = self.L.refb_calculate_allele_counts(self.d, group, C.c_char(a), n, _ptr(out, C.c_double))
        assert rows == n
        return out

    def optimal_haplotype(self, eff_id, na=b"-"):
        buf = C.create_string_buffer(self.n_markers + 1)
        self.L.refb_optimal_haplotype(self.d, eff_id, C.c_char(na), buf)
        return buf.raw[:self.n_markers]

    def optimal_possible_haplotype(self, group, eff_id, na=b"-"):
        buf = C.create_string_buffer(self.n_markers + 1)
        self.L.refb_optimal_possible_haplotype(self.d, group, eff_id, C.c_char(na), buf)
        return buf.raw[:self.n_markers]

    def optimal_bv(self, eff_id):
        self.L.refb_optimal_bv.restype = C.c_double
        return self.L.refb_optimal_bv(self.d, eff_id)

    def minimal_bv(self, eff_id):
        self.L.refb_minimal_bv.restype = C.c_double
        return self.L.refb_minimal_bv(self.d, eff_id)

    def optimal_possible_bv(self, group, eff_id):
        self.L.refb_optimal_possible_bv.restype = C.c_double
        return self.L.refb_optimal_possible_bv(self.d, group, eff_id)

    def split_by_bv(self, group, eff_id, top_n, low_is_best=False):
        return self.L.refb_split_by_bv(self.d, group, eff_id, top_n, int(low_is_best))

    def find_crossovers(self, input_file, output_file, window=1, certain=True):
        return self.L.refb_recombinations_from_file(self.d, str(input_file).encode(), str(output_file).encode(), int(window), int(bool(certain)))

    def split(self, kind, group, n=0, counts=None, probs=None):
        """the reference's group-split utilities (refb_split in ref_bridge.c lists the kinds); returns the new group numbers"""
        cap = max(self.group_size(group), n, 1)
        out = np.zeros(cap, dtype=np.uint32)
        c = _u32(counts if counts is not None else [0])
        pr = np.ascontiguousarray(probs if probs is not None else [0.0], dtype=np.float64)
        made = self.L.refb_split(self.d, kind, group, n, _ptr(c, U), pr.ctypes.data_as(C.POINTER(C.c_double)), cap, _ptr(out, U))
        return [int(x) for x in out[:min(made, cap)]]

    def make_group_from(self, indexes):
        a = _u32(indexes)
        return self.L.refb_make_group_from(self.d, len(a), _ptr(a, U))

    def combine_groups(self, groups):
        a = _u32(groups)
        return self.L.refb_combine_groups(self.d, len(a), _ptr(a, U))

    def delete_group(self, group):
        self.L.refb_delete_group(self.d, group)

    # ---- savers --------------------------------------------------------
    def save_genotypes(self, f, group=0, markers_as_rows=False):
        self.L.refb_save_genotypes(self.d, f.encode(), group, int(markers_as_rows))

    def save_bvs(self, f, group, eff_id):
        self.L.refb_save_bvs(self.d, f.encode(), group, eff_id)

    def save_allele_counts(self, f, group, allele, markers_as_rows=False):
        a = allele if isinstance(allele, bytes) else allele.encode()
        self.L.refb_save_allele_counts(self.d, f.encode(), group, C.c_char(a), int(markers_as_rows))

    def save_local_bvs(self, f, group, blocks, eff_id, headers=True):
        self.L.refb_save_local_bvs(self.d, f.encode(), group, blocks.h, eff_id, int(headers))

    def save_pedigrees(self, f, group=0, full=True):
        self.L.refb_save_pedigrees(self.d, f.encode(), group, int(full))


class RefBlocks:
    def __init__(self, sim, h):
        self.sim, self.h = sim, h

    @property
    def n_blocks(self):
        return self.sim.L.refb_blocks_n(self.h)

    def export(self):
        nb = self.n_blocks
        tot = self.sim.L.refb_blocks_total(self.h)
        off = np.zeros(nb + 1, dtype=np.uint32)
        mk = np.zeros(max(tot, 1), dtype=np.uint32)
        self.sim.L.refb_blocks_export(self.h, _ptr(off, U), _ptr(mk, U))
        return off, mk[:tot]

    def free(self):
        if self.h:
            self.sim.L.refb_blocks_free(self.h)
            self.h = None
