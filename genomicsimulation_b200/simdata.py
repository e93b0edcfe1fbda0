"""SimData — ctypes binding of the host-side mirror (include/gsb200_gsc.h).

Method names follow oracle/ref.py and oracle/port.py so the parity tests run one scenario
through the reference, the restatement and this engine unchanged.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GenOptions, MarkerBlocks, EngineError

DRAWS_HOST, DRAWS_NATIVE = 0, 1


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _opts(family_size=1, track_pedigree=True, allocate_ids=True, name_offspring=False,
          retain=True, name_prefix=None, file_prefix=None, save_pedigree=False, save_bvs=0, save_genotypes=False):
    return GenOptions(bool(name_offspring), name_prefix.encode() if name_prefix is not None else None,
                      family_size, bool(track_pedigree), bool(allocate_ids),
                      file_prefix.encode() if file_prefix is not None else None, bool(save_pedigree), int(save_bvs),
                      bool(save_genotypes), bool(retain))


class Blocks:
    """Owns a gsb_MarkerBlocks."""

    def __init__(self, L, mb, owned=True, keep=None):
        self.L, self.mb, self.owned, self.keep = L, mb, owned, keep

    @classmethod
    def from_csr(cls, L, offsets, markers):
        o = np.ascontiguousarray(offsets, dtype=np.uint32)
        m = np.ascontiguousarray(markers, dtype=np.uint32)
        mb = MarkerBlocks(len(o) - 1, o.ctypes.data_as(C.POINTER(C.c_uint)), m.ctypes.data_as(C.POINTER(C.c_uint)))
        return cls(L, mb, owned=False, keep=(o, m))

    def export(self):
        nb = self.mb.num_blocks
        if not self.mb.offsets:
            return np.zeros(1, dtype=np.uint32), np.zeros(0, dtype=np.uint32)
        off = np.ctypeslib.as_array(self.mb.offsets, shape=(nb + 1,)).astype(np.uint32).copy()
        mk = np.ctypeslib.as_array(self.mb.markers, shape=(max(int(off[nb]), 1),)).astype(np.uint32)[:off[nb]].copy()
        return off, mk

    def free(self):
        if self.owned and self.mb.offsets:
            self.L.gsb_delete_markerblocks(C.byref(self.mb))
        self.owned = False

    def __del__(self):
        self.free()


class SimData:
    def __init__(self, n_markers, device=0, quiet=True):
        self.L = _lib.host()
        self.E = _lib.engine()
        self.M = int(n_markers)
        self.d = self.L.gsb_create_simdata(device, self.M)
        if not self.d:
            raise EngineError("gsb_create_simdata failed: %s" % self.E.gsb200_last_error().decode())
        self.d = C.c_void_p(self.d)
        self.ctx = C.c_void_p(self.L.gsb_context(self.d))
        self._keep = []
        if quiet:
            self.L.gsb_set_print(self.d, None)

    @classmethod
    def from_dataset(cls, ds, device=0):
        s = cls(ds["n_markers"], device=device)
        s.founders = s.add_genotypes(ds["genotypes"], ds.get("names"))
        for m in ds["maps"]:
            s.add_map(m)
        for e in ds["effects"]:
            s.add_effects(e)
        if ds.get("marker_names"):
            s.set_marker_names(ds["marker_names"])
        return s

    def close(self):
        if self.d:
            self.L.gsb_delete_simdata(self.d)
            self.d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self):
        st = self.L.gsb_last_status(self.d)
        if st != 0:
            raise EngineError("engine status %d: %s" % (st, self.L.gsb_last_message(self.d).decode()))

    # ---- draws -------------------------------------------------------------
    def use_host_draws(self, unif_ptr, rpois_ptr):
        """All draws from a host source, in the reference's order (replay / R-compatible mode)."""
        self.L.gsb_set_draw_mode(self.d, DRAWS_HOST, 0)
        self.L.gsb_set_draw_source(self.d, unif_ptr, rpois_ptr)

    def use_native_draws(self, seed, unif_ptr=None, rpois_ptr=None):
        """Meiosis draws from the device Philox stream; parent picks from the host source."""
        self.L.gsb_set_draw_mode(self.d, DRAWS_NATIVE, seed)
        self.L.gsb_set_draw_source(self.d, unif_ptr, rpois_ptr)

    # ---- setup -------------------------------------------------------------
    def add_genotypes(self, g, names=None):
        g = np.ascontiguousarray(g, dtype=np.uint8)
        n = g.shape[0]
        arr = None
        if names is not None:
            arr = (C.c_char_p * n)(*[nm.encode() if nm is not None else None for nm in names])
        grp = self.L.gsb_load_genotypes(self.d, n, _p(g), arr)
        self._check()
        return grp

    def load_genotype_file(self, path):
        grp = self.L.gsb_load_genotype_file(self.d, str(path).encode())
        self._check()
        return grp

    def add_map(self, m):
        lam = np.ascontiguousarray(m["lam"], dtype=np.float64)
        off = np.ascontiguousarray(m["chr_offset"], dtype=np.uint32)
        dists = np.ascontiguousarray(m["dists"], dtype=np.float64)
        mix = np.ascontiguousarray(m["marker_index"], dtype=np.uint32)
        has = np.ascontiguousarray(m["has_dists"], dtype=np.uint8)
        mid = self.L.gsb_load_map(self.d, len(lam), _p(lam), _p(off), _p(dists), _p(mix), _p(has))
        self._check()
        return mid

    def add_effects(self, e):
        cumn = np.ascontiguousarray(e["cumn"], dtype=np.uint32)
        allele = np.ascontiguousarray(e["allele"], dtype=np.uint8)
        eff = np.ascontiguousarray(e["eff"], dtype=np.float64)
        centre = None if e.get("centre") is None else np.ascontiguousarray(e["centre"], dtype=np.float64)
        eid = self.L.gsb_load_effects(self.d, _p(cumn), _p(allele), _p(eff), _p(centre) if centre is not None else None)
        self._check()
        return eid

    def map_arg(self, map_index):
        return map_index + 1          # ids are index + 1 here

    def eff_arg(self, eff_index):
        return eff_index + 1

    # ---- inspection ----------------------------------------------------------
    @property
    def n_markers(self):
        return self.M

    @property
    def n_genotypes(self):
        return self.L.gsb_get_n_genotypes(self.d)

    @property
    def current_id(self):
        return self.L.gsb_get_current_id(self.d)

    def group_size(self, group):
        return self.L.gsb_get_group_size(self.d, group)

    def export_group(self, group=0, genotypes=True):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        g = np.zeros((n, 2 * self.M), dtype=np.uint8) if genotypes else None
        ids, p1, p2, gr, gi = (np.zeros(n, dtype=np.uint32) for _ in range(5))
        got = self.L.gsb_get_group_data(self.d, group, n, _p(g) if genotypes else None,
                                        _p(ids), _p(p1), _p(p2), _p(gr), _p(gi))
        self._check()
        assert got == n
        return dict(genotypes=g, ids=ids, ped1=p1, ped2=p2, groups=gr, indexes=gi)

    def group_indexes(self, group):
        return self.export_group(group, genotypes=False)["indexes"]

    def group_names(self, group):
        n = self.n_genotypes if group == 0 else self.group_size(group)
        out = []
        for i in range(n):
            nm = self.L.gsb_get_group_member_name(self.d, group, i)
            out.append(nm.decode() if nm is not None else None)
        return out

    # ---- crossing --------------------------------------------------------------
    def make_random_crosses(self, group, n, cap=0, map_id=0, **kw):
        g = self.L.gsb_make_random_crosses(self.d, group, n, cap, map_id, _opts(**kw))
        self._check()
        return g

    def make_random_crosses_between(self, g1, g2, n, cap1=0, cap2=0, map1=0, map2=0, **kw):
        g = self.L.gsb_make_random_crosses_between(self.d, g1, g2, n, cap1, cap2, map1, map2, _opts(**kw))
        self._check()
        return g

    def make_targeted_crosses(self, p1, p2, map1=0, map2=0, **kw):
        a = np.ascontiguousarray(p1, dtype=np.uint32)
        b = np.ascontiguousarray(p2, dtype=np.uint32)
        g = self.L.gsb_make_targeted_crosses(self.d, len(a), _p(a), _p(b), map1, map2, _opts(**kw))
        self._check()
        return g

    def make_crosses_from_file(self, path, map1=0, map2=0, **kw):
        g = self.L.gsb_make_crosses_from_file(self.d, path.encode(), map1, map2, _opts(**kw))
        self._check()
        return g

    def make_double_crosses_from_file(self, path, map1=0, map2=0, **kw):
        g = self.L.gsb_make_double_crosses_from_file(self.d, path.encode(), map1, map2, _opts(**kw))
        self._check()
        return g

    def make_all_unidirectional_crosses(self, group, map_id=0, **kw):
        g = self.L.gsb_make_all_unidirectional_crosses(self.d, group, map_id, _opts(**kw))
        self._check()
        return g

    def self_n_times(self, n, group, map_id=0, **kw):
        g = self.L.gsb_self_n_times(self.d, n, group, map_id, _opts(**kw))
        self._check()
        return g

    def make_doubled_haploids(self, group, map_id=0, **kw):
        g = self.L.gsb_make_doubled_haploids(self.d, group, map_id, _opts(**kw))
        self._check()
        return g

    def make_clones(self, group, inherit_names=False, **kw):
        g = self.L.gsb_make_clones(self.d, group, bool(inherit_names), _opts(**kw))
        self._check()
        return g

    # ---- evaluation ----------------------------------------------------------
    def _take(self, dm):
        self._check()
        n = dm.dim1 * dm.dim2
        out = np.ctypeslib.as_array(dm.data, shape=(n,)).copy().reshape(dm.dim1, dm.dim2) if n else np.zeros((dm.dim1, dm.dim2))
        self.L.gsb_delete_dmatrix(C.byref(dm))
        return out

    def calculate_bvs(self, group, eff_id):
        return self._take(self.L.gsb_calculate_bvs(self.d, group, eff_id)).reshape(-1)

    def get_group_bvs(self, group, eff_id, out=None):
        """gsc_get_group_bvs: GEBVs into a caller-owned array (what see.GEBVs does, r-wrappers.c:795-813)."""
        n = self.group_size(group)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        got = self.L.gsb_get_group_bvs(self.d, group, eff_id, min(n, len(out)), _p(out))
        self._check()
        return out[:got]

    def calculate_allele_counts(self, group, allele):
        a = allele if isinstance(allele, bytes) else allele.encode()
        return self._take(self.L.gsb_calculate_allele_counts(self.d, group, C.c_char(a)))

    def calculate_local_bvs(self, group, *args):
        """(group, blocks, eff_id) like oracle/ref.py, or (group, offsets, markers, eff_id) like oracle/port.py"""
        if len(args) == 3:
            blocks, eff_id = Blocks.from_csr(self.L, args[0], args[1]), args[2]
        else:
            blocks, eff_id = args
        return self._take(self.L.gsb_calculate_local_bvs(self.d, group, blocks.mb, eff_id))

    def blocks_evenlength(self, n, map_id=0, n_chr=None):
        return Blocks(self.L, self.L.gsb_create_evenlength_blocks_each_chr(self.d, map_id, n))

    def set_marker_names(self, names):
        arr = (C.c_char_p * self.M)(*[n if isinstance(n, bytes) else str(n).encode() for n in names])
        self.L.gsb_set_marker_names(self.d, arr)

    def blocks_from_file(self, path):
        return Blocks(self.L, self.L.gsb_load_blocks(self.d, path if isinstance(path, bytes) else str(path).encode()))

    def optimal_haplotype(self, eff_id, na=b"-"):
        buf = C.create_string_buffer(self.M + 1)
        self.L.gsb_calculate_optimal_haplotype(self.d, eff_id, C.c_char(na), buf)
        return buf.raw[:self.M]

    def optimal_bv(self, eff_id):
        return self.L.gsb_calculate_optimal_bv(self.d, eff_id)

    def minimal_bv(self, eff_id):
        return self.L.gsb_calculate_minimal_bv(self.d, eff_id)

    def optimal_possible_haplotype(self, group, eff_id, na=b"-"):
        buf = C.create_string_buffer(self.M + 1)
        self.L.gsb_calculate_optimal_possible_haplotype(self.d, group, eff_id, C.c_char(na), buf)
        self._check()
        return buf.raw[:self.M]

    def optimal_possible_bv(self, group, eff_id):
        v = self.L.gsb_calculate_optimal_possible_bv(self.d, group, eff_id)
        self._check()
        return v

    def split_by_bv(self, group, eff_id, top_n, low_is_best=False):
        g = self.L.gsb_split_by_bv(self.d, group, eff_id, top_n, bool(low_is_best))
        self._check()
        return g

    # ---- sharded recurrent selection (include/gsb200_gsc.h gsb_shard_*) -------------------
    def shard_init(self, rank, world, max_generation, max_pool):
        if self.L.gsb_shard_init(self.d, rank, world, max_generation, max_pool) != 0:
            self._check()

    def shard_handle(self):
        buf = C.create_string_buffer(_lib.SHARD_HANDLE_BYTES)
        if self.L.gsb_shard_handle(self.d, buf) != 0:
            self._check()
        return buf.raw

    def shard_connect(self, handles):
        if self.L.gsb_shard_connect(self.d, handles) != 0:
            self._check()

    def shard_select_and_cross(self, group, n_total, eff_id, top_n, n_offspring_total, low_is_best=False, map_id=0, **kw):
        g = self.L.gsb_shard_select_and_cross(self.d, group, n_total, eff_id, top_n, bool(low_is_best), n_offspring_total,
                                              map_id, _opts(**kw))
        self._check()
        return g

    def shard_finish(self):
        """waits for the generation in flight and reports its verdict (gsb_shard_finish)"""
        if self.L.gsb_shard_finish(self.d) != 0:
            self._check()

    def shard_local_bvs(self, out):
        got = self.L.gsb_shard_get_local_bvs(self.d, len(out), _p(out))
        self._check()
        return out[:got]

    def shard_last_selection(self, n_total, top_n, bvs=True):
        bv = np.empty(n_total, dtype=np.float64) if bvs else None
        sel = np.empty(top_n, dtype=np.uint32)
        got = self.L.gsb_shard_last_selection(self.d, n_total, _p(bv) if bvs else None, top_n, _p(sel))
        self._check()
        assert got == top_n
        return sel, bv

    # ---- savers (same method names as oracle/ref.py) -------------------------------------
    def save_genotypes(self, f, group=0, markers_as_rows=False):
        self.L.gsb_save_genotypes(str(f).encode(), self.d, group, bool(markers_as_rows))
        self._check()

    def save_bvs(self, f, group, eff_id):
        self.L.gsb_save_bvs(str(f).encode(), self.d, group, eff_id)
        self._check()

    def save_allele_counts(self, f, group, allele, markers_as_rows=False):
        a = allele if isinstance(allele, bytes) else allele.encode()
        self.L.gsb_save_allele_counts(str(f).encode(), self.d, group, C.c_char(a), bool(markers_as_rows))
        self._check()

    def save_local_bvs(self, f, group, blocks, eff_id, headers=True):
        self.L.gsb_save_local_bvs(str(f).encode(), self.d, group, blocks.mb, eff_id, bool(headers))
        self._check()

    def save_pedigrees(self, f, group=0, full=True):
        self.L.gsb_save_pedigrees(str(f).encode(), self.d, group, bool(full))
        self._check()

    def make_group_from(self, indexes):
        a = np.ascontiguousarray(indexes, dtype=np.uint32)
        return self.L.gsb_make_group_from(self.d, len(a), _p(a))

    def delete_group(self, group):
        self.L.gsb_delete_group(self.d, group)

    def combine_groups(self, groups):
        a = np.ascontiguousarray(groups, dtype=np.uint32)
        return self.L.gsb_combine_groups(self.d, len(a), _p(a))

    # ---- find.crossovers (gsb_calculate_recombinations_from_file, core.c:7760-7818)
    def find_crossovers(self, input_file, output_file, window=1, certain=True):
        st = self.L.gsb_calculate_recombinations_from_file(self.d, str(input_file).encode(), str(output_file).encode(), int(window), int(bool(certain)))
        self._check()
        return st

    def min_recombinations(self, parent1, p1num, parent2, p2num, offspring, window=1, certain=True, map_id=0):
        out = np.zeros(self.M, dtype=np.int32)
        self.L.gsb_calculate_min_recombinations(self.d, map_id, parent1, p1num, parent2, p2num, offspring, int(window), int(bool(certain)), _p(out))
        self._check()
        return out

    # ---- group-split utilities (gsb_split_*, core.c:2980-3784): each returns the list of new group numbers
    def _split(self, fn, group, cap, *args):
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        n = fn(self.d, group, *args, _p(out))
        return [int(x) for x in out[:min(n, cap)]]

    def split_into_families(self, group):
        cap = self.group_size(group)
        return self._split(self.L.gsb_split_into_families, group, cap, cap)

    def split_into_halfsib_families(self, group, parent):
        cap = self.group_size(group)
        return self._split(self.L.gsb_split_into_halfsib_families, group, cap, parent, cap)

    def split_into_individuals(self, group):
        cap = self.group_size(group)
        return self._split(self.L.gsb_split_into_individuals, group, cap, cap)

    def split_evenly_into_two(self, group):
        return self.L.gsb_split_evenly_into_two(self.d, group)

    def split_evenly_into_n(self, group, n):
        return self._split(self.L.gsb_split_evenly_into_n, group, n, n)

    def split_into_buckets(self, group, counts):
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        return self._split(self.L.gsb_split_into_buckets, group, len(c) + 1, len(c) + 1, _p(c))

    def split_randomly_into_two(self, group):
        return self.L.gsb_split_randomly_into_two(self.d, group)

    def split_randomly_into_n(self, group, n):
        return self._split(self.L.gsb_split_randomly_into_n, group, n, n)

    def split_by_probabilities(self, group, probs):
        pr = np.ascontiguousarray(probs, dtype=np.float64)
        return self._split(self.L.gsb_split_by_probabilities, group, len(pr) + 1, len(pr) + 1, _p(pr))
