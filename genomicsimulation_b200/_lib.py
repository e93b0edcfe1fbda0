"""ctypes loading of the two product libraries. Fails loudly when they are missing."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ENGINE = os.path.join(_HERE, "libgsb200.so")
LIB_HOST = os.path.join(_HERE, "libgsb200_gsc.so")

U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
VP = C.c_void_p


class EngineError(RuntimeError):
    pass


class Draws(C.Structure):
    """gsb200_draws (include/gsb200.h)."""
    _fields_ = [("mode", U32), ("seed", U64), ("stream", U64), ("first_unit", U64),
                ("n_entries", U64), ("n_crossovers", VP), ("start_strand", VP),
                ("n_positions", U64), ("positions", VP)]


class GenOptions(C.Structure):
    """gsb_GenOptions == gsc_GenOptions field for field (sim-operations.h:629-669)."""
    _fields_ = [("will_name_offspring", C.c_bool), ("offspring_name_prefix", C.c_char_p),
                ("family_size", C.c_uint), ("will_track_pedigree", C.c_bool), ("will_allocate_ids", C.c_bool),
                ("filename_prefix", C.c_char_p), ("will_save_pedigree_to_file", C.c_bool),
                ("will_save_bvs_to_file", C.c_uint), ("will_save_alleles_to_file", C.c_bool),
                ("will_save_to_simdata", C.c_bool)]


class DecimalMatrix(C.Structure):
    _fields_ = [("dim1", C.c_uint), ("dim2", C.c_uint), ("data", C.POINTER(C.c_double))]


class MarkerBlocks(C.Structure):
    _fields_ = [("num_blocks", C.c_uint), ("offsets", C.POINTER(C.c_uint)), ("markers", C.POINTER(C.c_uint))]


def build():
    """Compile both libraries in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


_engine = None
_host = None

ENGINE_SYMBOLS = {
    # name: (restype, argtypes)
    "gsb200_abi_version": (I32, []),
    "gsb200_last_error": (C.c_char_p, []),
    "gsb200_create": (I32, [C.POINTER(VP), I32, U32]),
    "gsb200_destroy": (None, [VP]),
    "gsb200_synchronize": (I32, [VP]),
    "gsb200_stream": (VP, [VP]),
    "gsb200_store_reserve": (I32, [VP, U64]),
    "gsb200_store_capacity": (U64, [VP]),
    "gsb200_store_planes": (U32, [VP]),
    "gsb200_store_plane_words": (U32, [VP]),
    "gsb200_store_row_bytes": (U64, [VP]),
    "gsb200_store_device_ptr": (VP, [VP]),
    "gsb200_store_upload_chars": (I32, [VP, U32, VP, VP]),
    "gsb200_store_download_chars": (I32, [VP, U32, VP, VP]),
    "gsb200_store_download_chars_range": (I32, [VP, U32, VP, U32, U32, VP]),
    "gsb200_store_gather_rows_dev": (I32, [VP, U32, VP, VP]),
    "gsb200_store_scatter_rows_dev": (I32, [VP, U32, VP, VP]),
    "gsb200_store_change_symbol": (I32, [VP, U32, C.c_char, C.c_char]),
    "gsb200_store_dictionary": (I32, [VP, VP, VP]),
    "gsb200_store_symbols": (I32, [VP, U32, VP, U32]),
    "gsb200_map_set": (I32, [VP, U32, U32, VP, VP, VP, VP, VP]),
    "gsb200_map_delete": (I32, [VP, U32]),
    "gsb200_map_n_chr": (U32, [VP, U32]),
    "gsb200_effects_set": (I32, [VP, U32, VP, VP, VP, VP]),
    "gsb200_effects_delete": (I32, [VP, U32]),
    "gsb200_cross": (I32, [VP, U32, VP, VP, U32, U32, VP, C.POINTER(Draws)]),
    "gsb200_cross_random_pairs": (I32, [VP, U32, U32, VP, U32, VP, U32, U32, U32, VP, U32, C.POINTER(Draws), VP, VP]),
    "gsb200_cross_random_pairs_begin": (I32, [VP, U32, U32, VP, U32, VP, U32, U32, U32, VP, U32, C.POINTER(Draws), VP, VP]),
    "gsb200_cross_random_pairs_end": (I32, [VP]),
    "gsb200_cross_dev": (I32, [VP, U32, VP, VP, U32, U32, VP, C.POINTER(Draws)]),
    "gsb200_cross_gebv_dev": (I32, [VP, U32, VP, VP, U32, U32, VP, C.POINTER(Draws), U32, VP, U32, VP]),
    "gsb200_cross_dev_status": (I32, [VP]),
    "gsb200_self_n_times": (I32, [VP, U32, VP, U32, U32, VP, C.POINTER(Draws)]),
    "gsb200_doubled_haploid": (I32, [VP, U32, VP, U32, VP, C.POINTER(Draws)]),
    "gsb200_clone": (I32, [VP, U32, VP, VP]),
    "gsb200_export_draws": (I32, [VP, U32, U64, U32, C.POINTER(Draws), VP, VP, VP, U64, C.POINTER(U64)]),
    "gsb200_gebv": (I32, [VP, U32, VP, U32, VP]),
    "gsb200_gebv_dev": (I32, [VP, U32, VP, U32, VP]),
    "gsb200_local_gebv": (I32, [VP, U32, VP, U32, VP, VP, U32, VP]),
    "gsb200_local_gebv_multi": (I32, [VP, U32, VP, U32, VP, VP, U32, VP, VP]),
    "gsb200_local_gebv_multi_dev": (I32, [VP, U32, VP, U32, VP, VP, U32, VP, VP]),
    "gsb200_allele_counts": (I32, [VP, U32, VP, C.c_char, I32, VP]),
    "gsb200_allele_presence": (I32, [VP, U32, VP, VP]),
    "gsb200_top_n": (I32, [VP, U32, VP, U32, I32, VP]),
    "gsb200_kernel_launches": (U64, [VP]),
    "gsb200_shard_range": (None, [U64, U32, U32, C.POINTER(U64), C.POINTER(U64)]),
    "gsb200_shard_create": (I32, [VP, U32, U32, U64, U32, C.POINTER(VP)]),
    "gsb200_shard_destroy": (None, [VP]),
    "gsb200_shard_handle": (I32, [VP, VP]),
    "gsb200_shard_connect": (I32, [VP, VP]),
    "gsb200_shard_select": (I32, [VP, U32, VP, VP, U64, U32, U32, I32]),
    "gsb200_shard_select_run": (I32, [VP, U32, VP, U32, VP, U64, U32, U32, I32]),
    "gsb200_shard_select_dev": (I32, [VP, U32, VP, VP, U64, U32, U32, I32]),
    "gsb200_shard_cross": (I32, [VP, U32, U32, VP, C.POINTER(Draws), VP, VP]),
    "gsb200_shard_cross_begin": (I32, [VP, U32, U32, VP, C.POINTER(Draws), I32]),
    "gsb200_shard_cross_begin_run": (I32, [VP, U32, U32, VP, U32, C.POINTER(Draws), I32]),
    "gsb200_shard_cross_picks": (I32, [VP, VP, VP, VP]),
    "gsb200_shard_finish": (I32, [VP]),
    "gsb200_shard_poll": (I32, [VP]),
    "gsb200_shard_cross_dev": (I32, [VP, U32, U32, VP, U32, C.POINTER(Draws), VP]),
    "gsb200_shard_selected": (I32, [VP, VP, VP, VP]),
    "gsb200_shard_local_bvs": (I32, [VP, U32, VP]),
    "gsb200_shard_generation_size": (U64, [VP]),
    "gsb200_shard_pool_size": (U32, [VP]),
    "gsb200_shard_pool_device_ptr": (VP, [VP]),
    "gsb200_shard_set_timing": (I32, [VP, I32]),
    "gsb200_shard_epoch": (U32, [VP]),
    "gsb200_shard_phase_ms": (I32, [VP, U32, VP]),
    "gsb200_shard_status": (I32, [VP]),
}
SHARD_HANDLE_BYTES = 128

UNIF_FN = C.CFUNCTYPE(C.c_double)
RPOIS_FN = C.CFUNCTYPE(C.c_double, C.c_double)
PRINT_FN = C.CFUNCTYPE(None, C.c_char_p)
UINT = C.c_uint

HOST_SYMBOLS = {
    "gsb_create_simdata": (VP, [I32, UINT]),
    "gsb_delete_simdata": (None, [VP]),
    "gsb_context": (VP, [VP]),
    "gsb_last_status": (I32, [VP]),
    "gsb_last_message": (C.c_char_p, [VP]),
    "gsb_set_print": (None, [VP, VP]),
    "gsb_set_draw_mode": (None, [VP, I32, U64]),
    "gsb_set_draw_source": (None, [VP, VP, VP]),
    "gsb_load_genotypes": (UINT, [VP, UINT, VP, VP]),
    "gsb_load_genotype_file": (UINT, [VP, C.c_char_p]),
    "gsb_load_map": (UINT, [VP, UINT, VP, VP, VP, VP, VP]),
    "gsb_load_effects": (UINT, [VP, VP, VP, VP, VP]),
    "gsb_delete_recombination_map": (None, [VP, UINT]),
    "gsb_delete_eff_set": (None, [VP, UINT]),
    "gsb_get_n_genotypes": (UINT, [VP]),
    "gsb_get_n_markers": (UINT, [VP]),
    "gsb_get_marker_names": (VP, [VP]),
    "gsb_get_map_ids": (UINT, [VP, UINT, VP]),
    "gsb_get_eff_set_ids": (UINT, [VP, UINT, VP]),
    "gsb_get_index_of_name": (UINT, [VP, C.c_char_p]),
    "gsb_get_current_id": (UINT, [VP]),
    "gsb_get_group_size": (UINT, [VP, UINT]),
    "gsb_get_group_data": (UINT, [VP, UINT, UINT, VP, VP, VP, VP, VP, VP]),
    "gsb_get_group_member_name": (C.c_char_p, [VP, UINT, UINT]),
    "gsb_make_group_from": (UINT, [VP, UINT, VP]),
    "gsb_delete_group": (None, [VP, UINT]),
    "gsb_combine_groups": (UINT, [VP, UINT, VP]),
    "gsb_calculate_min_recombinations": (C.c_int, [VP, UINT, UINT, UINT, UINT, UINT, UINT, C.c_int, C.c_int, VP]),
    "gsb_calculate_recombinations_from_file": (C.c_int, [VP, C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
    "gsb_split_into_families": (UINT, [VP, UINT, UINT, VP]),
    "gsb_split_into_halfsib_families": (UINT, [VP, UINT, C.c_int, UINT, VP]),
    "gsb_split_into_individuals": (UINT, [VP, UINT, UINT, VP]),
    "gsb_split_evenly_into_two": (UINT, [VP, UINT]),
    "gsb_split_evenly_into_n": (UINT, [VP, UINT, UINT, VP]),
    "gsb_split_into_buckets": (UINT, [VP, UINT, UINT, VP, VP]),
    "gsb_split_randomly_into_two": (UINT, [VP, UINT]),
    "gsb_split_randomly_into_n": (UINT, [VP, UINT, UINT, VP]),
    "gsb_split_by_probabilities": (UINT, [VP, UINT, UINT, VP, VP]),
    "gsb_make_random_crosses": (UINT, [VP, UINT, UINT, UINT, UINT, GenOptions]),
    "gsb_make_random_crosses_between": (UINT, [VP, UINT, UINT, UINT, UINT, UINT, UINT, UINT, GenOptions]),
    "gsb_make_targeted_crosses": (UINT, [VP, UINT, VP, VP, UINT, UINT, GenOptions]),
    "gsb_self_n_times": (UINT, [VP, UINT, UINT, UINT, GenOptions]),
    "gsb_make_doubled_haploids": (UINT, [VP, UINT, UINT, GenOptions]),
    "gsb_make_clones": (UINT, [VP, UINT, C.c_bool, GenOptions]),
    "gsb_make_all_unidirectional_crosses": (UINT, [VP, UINT, UINT, GenOptions]),
    "gsb_make_crosses_from_file": (UINT, [VP, C.c_char_p, UINT, UINT, GenOptions]),
    "gsb_make_double_crosses_from_file": (UINT, [VP, C.c_char_p, UINT, UINT, GenOptions]),
    "gsb_calculate_bvs": (DecimalMatrix, [VP, UINT, UINT]),
    "gsb_get_group_bvs": (UINT, [VP, UINT, UINT, UINT, VP]),
    "gsb_calculate_allele_counts": (DecimalMatrix, [VP, UINT, C.c_char]),
    "gsb_calculate_local_bvs": (DecimalMatrix, [VP, UINT, MarkerBlocks, UINT]),
    "gsb_create_evenlength_blocks_each_chr": (MarkerBlocks, [VP, UINT, UINT]),
    "gsb_set_marker_names": (None, [VP, VP]),
    "gsb_load_blocks": (MarkerBlocks, [VP, C.c_char_p]),
    "gsb_calculate_optimal_haplotype": (None, [VP, UINT, C.c_char, VP]),
    "gsb_calculate_optimal_bv": (C.c_double, [VP, UINT]),
    "gsb_calculate_minimal_bv": (C.c_double, [VP, UINT]),
    "gsb_calculate_optimal_possible_haplotype": (None, [VP, UINT, UINT, C.c_char, VP]),
    "gsb_calculate_optimal_possible_bv": (C.c_double, [VP, UINT, UINT]),
    "gsb_split_by_bv": (UINT, [VP, UINT, UINT, UINT, C.c_bool]),
    "gsb_delete_dmatrix": (None, [C.POINTER(DecimalMatrix)]),
    "gsb_delete_markerblocks": (None, [C.POINTER(MarkerBlocks)]),
    "gsb_save_genotypes": (None, [C.c_char_p, VP, UINT, C.c_bool]),
    "gsb_save_allele_counts": (None, [C.c_char_p, VP, UINT, C.c_char, C.c_bool]),
    "gsb_save_pedigrees": (None, [C.c_char_p, VP, UINT, C.c_bool]),
    "gsb_save_bvs": (None, [C.c_char_p, VP, UINT, UINT]),
    "gsb_save_local_bvs": (None, [C.c_char_p, VP, UINT, MarkerBlocks, UINT, C.c_bool]),
    "gsb_shard_init": (I32, [VP, UINT, UINT, C.c_ulonglong, UINT]),
    "gsb_shard_handle": (I32, [VP, VP]),
    "gsb_shard_connect": (I32, [VP, VP]),
    "gsb_shard_select_and_cross": (UINT, [VP, UINT, C.c_ulonglong, UINT, UINT, C.c_bool, C.c_ulonglong, UINT, GenOptions]),
    "gsb_shard_finish": (I32, [VP]),
    "gsb_shard_get_local_bvs": (UINT, [VP, UINT, VP]),
    "gsb_shard_last_selection": (UINT, [VP, C.c_ulonglong, VP, UINT, VP]),
}


def _load(path, table):
    if not os.path.exists(path):
        raise EngineError("%s is not built; run `make -C genomicsimulation_b200` (there is no CPU fallback)" % path)
    # RTLD_LOCAL: libgsb200_gsc.so exports the reference's own gsc_* names (include/gsb200_compat.h); in a process that also
    # loads the reference itself (the parity tests) they must not interpose each other
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (res, args) in table.items():
        fn = getattr(lib, name)          # AttributeError if the C-ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def engine():
    """libgsb200.so with prototypes set."""
    global _engine
    if _engine is None:
        _engine = _load(LIB_ENGINE, ENGINE_SYMBOLS)
    return _engine


def host():
    """libgsb200_gsc.so with prototypes set (loads the engine first)."""
    global _host
    if _host is None:
        engine()
        _host = _load(LIB_HOST, HOST_SYMBOLS)
    return _host
