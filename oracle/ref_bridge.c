/* oracle/ref_bridge.c — TEST INFRASTRUCTURE ONLY.
 *
 * A flat C interface over the UNMODIFIED reference core, so that tests and the
 * CPU-baseline leg of bench.py can drive it through ctypes without having to
 * mirror the reference's struct layouts in Python.  This file is compiled into
 * oracle/_ref/libgsref.so together with /root/reference/src/sim-operations.c
 * (from where it lies, never copied) and the R shim.  It only *calls* the
 * reference's public API (sim-operations.h); it contains no algorithm.
 */
#include "sim-operations.h"
#include "draws.h"

#define EXPORT __attribute__((visibility("default")))

/* ---------------- lifetime and loading ---------------- */

EXPORT void* refb_new(void) { return gsc_create_empty_simdata(); }
EXPORT void  refb_free(void* d) { gsc_delete_simdata((gsc_SimData*) d); }

/* ids[3] = { group, map id, effect-set id } (core.c:7453-7496) */
EXPORT void refb_load(void* d, const char* geno, const char* map, const char* eff, unsigned ids[3]) {
    struct gsc_MultiIDSet s = gsc_load_data_files((gsc_SimData*) d, geno, map, eff, GSC_DETECT_FILE_FORMAT);
    ids[0] = s.group.num; ids[1] = s.map.id; ids[2] = s.effSet.id;
}
EXPORT unsigned refb_load_effects(void* d, const char* f) { return gsc_load_effectfile((gsc_SimData*) d, f).id; }
EXPORT unsigned refb_load_map(void* d, const char* f) { return gsc_load_mapfile((gsc_SimData*) d, f).id; }
EXPORT unsigned refb_load_genotypes(void* d, const char* f) {
    return gsc_load_genotypefile((gsc_SimData*) d, f, GSC_DETECT_FILE_FORMAT).num;
}

/* ---------------- inspection ---------------- */

EXPORT unsigned refb_n_markers(void* d) { return ((gsc_SimData*) d)->genome.n_markers; }
EXPORT unsigned refb_n_maps(void* d) { return ((gsc_SimData*) d)->genome.n_maps; }
EXPORT unsigned refb_n_eff_sets(void* d) { return ((gsc_SimData*) d)->n_eff_sets; }
EXPORT unsigned refb_current_id(void* d) { return ((gsc_SimData*) d)->current_id.id; }
EXPORT const char* refb_marker_name(void* d, unsigned m) { return ((gsc_SimData*) d)->genome.marker_names[m]; }

EXPORT unsigned refb_n_genotypes(void* d) {
    unsigned n = 0;
    for (gsc_AlleleMatrix* m = ((gsc_SimData*) d)->m; m != NULL; m = m->next) n += m->n_genotypes;
    return n;
}
EXPORT unsigned refb_group_size(void* d, unsigned group) {
    return gsc_get_group_size((gsc_SimData*) d, (gsc_GroupNum){.num = group});
}

/* Walk a group (0 = everything) in storage order; any output pointer may be NULL.
 * genotypes receives 2*n_markers chars per member (the reference's own layout). */
EXPORT unsigned refb_export_group(void* dv, unsigned group, unsigned cap,
                                  char* genotypes, unsigned* ids, unsigned* ped1, unsigned* ped2,
                                  unsigned* groups) {
    gsc_SimData* d = (gsc_SimData*) dv;
    size_t rowlen = 2 * (size_t) d->genome.n_markers;
    gsc_BidirectionalIterator it = gsc_create_bidirectional_iter(d, (gsc_GroupNum){.num = group});
    gsc_GenoLocation loc = gsc_set_bidirectional_iter_to_start(&it);
    unsigned n = 0;
    while (GSC_IS_VALID_LOCATION(loc) && n < cap) {
        if (genotypes) memcpy(genotypes + rowlen * n, gsc_get_alleles(loc), rowlen);
        if (ids) ids[n] = gsc_get_id(loc).id;
        if (ped1) ped1[n] = gsc_get_first_parent(loc).id;
        if (ped2) ped2[n] = gsc_get_second_parent(loc).id;
        if (groups) groups[n] = gsc_get_group(loc).num;
        ++n;
        loc = gsc_next_forwards(&it);
    }
    gsc_delete_bidirectional_iter(&it);
    return n;
}
EXPORT unsigned refb_group_indexes(void* d, unsigned group, unsigned cap, unsigned* out) {
    return gsc_get_group_indexes((gsc_SimData*) d, (gsc_GroupNum){.num = group}, cap, out);
}
/* name of the i-th member of a group in storage order, or NULL */
EXPORT const char* refb_group_member_name(void* dv, unsigned group, unsigned i) {
    gsc_SimData* d = (gsc_SimData*) dv;
    gsc_BidirectionalIterator it = gsc_create_bidirectional_iter(d, (gsc_GroupNum){.num = group});
    gsc_GenoLocation loc = gsc_set_bidirectional_iter_to_start(&it);
    while (GSC_IS_VALID_LOCATION(loc) && i > 0) { loc = gsc_next_forwards(&it); --i; }
    const char* nm = GSC_IS_VALID_LOCATION(loc) ? gsc_get_name(loc) : NULL;
    gsc_delete_bidirectional_iter(&it);
    return nm;
}

/* recombination map `map_index` (index, not id) */
EXPORT unsigned refb_map_id(void* d, unsigned map_index) { return ((gsc_SimData*) d)->genome.map_ids[map_index].id; }
EXPORT unsigned refb_map_n_chr(void* d, unsigned map_index) { return ((gsc_SimData*) d)->genome.maps[map_index].n_chr; }
/* info[0] = type (0 simple, 1 reorder), info[1] = n_markers, info[2] = first_marker_index (simple only) */
EXPORT double refb_map_chr_info(void* d, unsigned map_index, unsigned chr, unsigned info[3]) {
    gsc_LinkageGroup* lg = &((gsc_SimData*) d)->genome.maps[map_index].chrs[chr];
    if (lg->type == GSC_LINKAGEGROUP_SIMPLE) {
        info[0] = 0; info[1] = lg->map.simple.n_markers; info[2] = lg->map.simple.first_marker_index;
        return lg->map.simple.expected_n_crossovers;
    }
    info[0] = 1; info[1] = lg->map.reorder.n_markers; info[2] = 0;
    return lg->map.reorder.expected_n_crossovers;
}
/* dists gets n_markers doubles (raw bits preserved); marker_indexes the per-position marker index */
EXPORT void refb_map_chr_arrays(void* d, unsigned map_index, unsigned chr, double* dists, unsigned* marker_indexes) {
    gsc_LinkageGroup* lg = &((gsc_SimData*) d)->genome.maps[map_index].chrs[chr];
    if (lg->type == GSC_LINKAGEGROUP_SIMPLE) {
        for (unsigned i = 0; i < lg->map.simple.n_markers; ++i) {
            if (dists && lg->map.simple.dists) dists[i] = lg->map.simple.dists[i];
            if (marker_indexes) marker_indexes[i] = lg->map.simple.first_marker_index + i;
        }
    } else {
        for (unsigned i = 0; i < lg->map.reorder.n_markers; ++i) {
            if (dists && lg->map.reorder.dists) dists[i] = lg->map.reorder.dists[i];
            if (marker_indexes) marker_indexes[i] = lg->map.reorder.marker_indexes[i];
        }
    }
}
EXPORT int refb_map_chr_has_dists(void* d, unsigned map_index, unsigned chr) {
    gsc_LinkageGroup* lg = &((gsc_SimData*) d)->genome.maps[map_index].chrs[chr];
    return (lg->type == GSC_LINKAGEGROUP_SIMPLE ? lg->map.simple.dists : lg->map.reorder.dists) != NULL;
}

/* effect set `eff_index` (index, not id): CSR over markers (core.h:826-846) */
EXPORT unsigned refb_eff_id(void* d, unsigned eff_index) { return ((gsc_SimData*) d)->eff_set_ids[eff_index].id; }
EXPORT unsigned refb_eff_nnz(void* d, unsigned eff_index) {
    gsc_MarkerEffects* e = &((gsc_SimData*) d)->e[eff_index];
    return e->n_markers ? e->cumn_alleles[e->n_markers - 1] : 0;
}
EXPORT int refb_eff_export(void* d, unsigned eff_index, unsigned* cumn, char* allele, double* eff, double* centre) {
    gsc_MarkerEffects* e = &((gsc_SimData*) d)->e[eff_index];
    unsigned nnz = refb_eff_nnz(d, eff_index);
    memcpy(cumn, e->cumn_alleles, sizeof(unsigned) * e->n_markers);
    memcpy(allele, e->allele, nnz);
    memcpy(eff, e->eff, sizeof(double) * nnz);
    if (e->centre != NULL && centre != NULL) memcpy(centre, e->centre, sizeof(double) * e->n_markers);
    return e->centre != NULL;
}
EXPORT int refb_set_centres(void* d, unsigned eff_id, unsigned n, const double* values) {
    return gsc_change_eff_set_centres_to_values((gsc_SimData*) d, (gsc_EffectID){.id = eff_id}, n, values);
}

/* ---------------- crossing ---------------- */

/* opt[] = { family_size, track_pedigree, allocate_ids, name_offspring, retain,
 *           save_pedigree, save_bvs(effect id or 0), save_genotypes }   (core.h:629-669) */
static gsc_GenOptions make_opts(const unsigned opt[8], const char* name_prefix, const char* file_prefix) {
    gsc_GenOptions g = GSC_BASIC_OPT;
    g.family_size = opt[0];
    g.will_track_pedigree = opt[1] != 0;
    g.will_allocate_ids = opt[2] != 0;
    g.will_name_offspring = opt[3] != 0;
    g.offspring_name_prefix = name_prefix;
    g.will_save_to_simdata = opt[4] != 0;
    g.will_save_pedigree_to_file = opt[5] != 0;
    g.will_save_bvs_to_file = (gsc_EffectID){.id = opt[6]};
    g.will_save_alleles_to_file = opt[7] != 0;
    g.filename_prefix = file_prefix;
    return g;
}

EXPORT unsigned refb_make_random_crosses(void* d, unsigned group, unsigned n, unsigned cap, unsigned map_id,
                                         const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_random_crosses((gsc_SimData*) d, (gsc_GroupNum){.num = group}, n, cap,
                                   (gsc_MapID){.id = map_id}, make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_make_random_crosses_between(void* d, unsigned g1, unsigned g2, unsigned n,
                                                 unsigned cap1, unsigned cap2, unsigned map1, unsigned map2,
                                                 const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_random_crosses_between((gsc_SimData*) d, (gsc_GroupNum){.num = g1}, (gsc_GroupNum){.num = g2},
                                           n, cap1, cap2, (gsc_MapID){.id = map1}, (gsc_MapID){.id = map2},
                                           make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_make_targeted_crosses(void* d, unsigned n, const unsigned* p1, const unsigned* p2,
                                           unsigned map1, unsigned map2,
                                           const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_targeted_crosses((gsc_SimData*) d, n, p1, p2, (gsc_MapID){.id = map1},
                                     (gsc_MapID){.id = map2}, make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_make_all_unidirectional_crosses(void* d, unsigned group, unsigned map_id,
                                                     const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_all_unidirectional_crosses((gsc_SimData*) d, (gsc_GroupNum){.num = group},
                                               (gsc_MapID){.id = map_id}, make_opts(opt, np, fp)).num;
}
EXPORT void refb_optimal_haplotype(void* d, unsigned eff_id, char na, char* out) {
    gsc_calculate_optimal_haplotype((gsc_SimData*) d, (gsc_EffectID){.id = eff_id}, na, out);
}
EXPORT void refb_optimal_possible_haplotype(void* d, unsigned group, unsigned eff_id, char na, char* out) {
    gsc_calculate_optimal_possible_haplotype((gsc_SimData*) d, (gsc_GroupNum){.num = group}, (gsc_EffectID){.id = eff_id}, na, out);
}
EXPORT double refb_optimal_bv(void* d, unsigned eff_id) { return gsc_calculate_optimal_bv((gsc_SimData*) d, (gsc_EffectID){.id = eff_id}); }
EXPORT double refb_minimal_bv(void* d, unsigned eff_id) { return gsc_calculate_minimal_bv((gsc_SimData*) d, (gsc_EffectID){.id = eff_id}); }
EXPORT double refb_optimal_possible_bv(void* d, unsigned group, unsigned eff_id) {
    return gsc_calculate_optimal_possible_bv((gsc_SimData*) d, (gsc_GroupNum){.num = group}, (gsc_EffectID){.id = eff_id});
}
EXPORT unsigned refb_make_crosses_from_file(void* d, const char* path, unsigned map1, unsigned map2,
                                            const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_crosses_from_file((gsc_SimData*) d, path, (gsc_MapID){.id = map1}, (gsc_MapID){.id = map2},
                                      make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_make_double_crosses_from_file(void* d, const char* path, unsigned map1, unsigned map2,
                                                   const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_double_crosses_from_file((gsc_SimData*) d, path, (gsc_MapID){.id = map1}, (gsc_MapID){.id = map2},
                                             make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_self_n_times(void* d, unsigned n, unsigned group, unsigned map_id,
                                  const unsigned opt[8], const char* np, const char* fp) {
    return gsc_self_n_times((gsc_SimData*) d, n, (gsc_GroupNum){.num = group}, (gsc_MapID){.id = map_id},
                            make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_make_doubled_haploids(void* d, unsigned group, unsigned map_id,
                                           const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_doubled_haploids((gsc_SimData*) d, (gsc_GroupNum){.num = group}, (gsc_MapID){.id = map_id},
                                     make_opts(opt, np, fp)).num;
}
EXPORT unsigned refb_make_clones(void* d, unsigned group, int inherit_names,
                                 const unsigned opt[8], const char* np, const char* fp) {
    return gsc_make_clones((gsc_SimData*) d, (gsc_GroupNum){.num = group}, inherit_names != 0,
                           make_opts(opt, np, fp)).num;
}
/* one gamete, written at stride 2 into out (core.c:7850-7928) */
EXPORT void refb_generate_gamete(void* d, const char* parent, char* out, unsigned map_index) {
    gsc_generate_gamete((gsc_SimData*) d, parent, out, map_index);
}
EXPORT void refb_generate_doubled_haploid(void* d, const char* parent, char* out, unsigned map_index) {
    gsc_generate_doubled_haploid((gsc_SimData*) d, parent, out, map_index);
}

/* ---------------- evaluation ---------------- */

/* returns number of values written (<= cap) or the needed count when out == NULL */
EXPORT unsigned refb_calculate_bvs(void* d, unsigned group, unsigned eff_id, unsigned cap, double* out) {
    gsc_DecimalMatrix m = gsc_calculate_bvs((gsc_SimData*) d, (gsc_GroupNum){.num = group}, (gsc_EffectID){.id = eff_id});
    unsigned n = m.dim2;
    if (out) for (unsigned i = 0; i < n && i < cap; ++i) out[i] = m.matrix[0][i];
    gsc_delete_dmatrix(&m);
    return n;
}

EXPORT void* refb_blocks_evenlength(void* d, unsigned map_id, unsigned n) {
    gsc_MarkerBlocks* b = malloc(sizeof *b);
    *b = gsc_create_evenlength_blocks_each_chr((gsc_SimData*) d, (gsc_MapID){.id = map_id}, n);
    return b;
}
EXPORT void* refb_blocks_from_file(void* d, const char* file) {
    gsc_MarkerBlocks* b = malloc(sizeof *b);
    *b = gsc_load_blocks((gsc_SimData*) d, file);
    return b;
}
/* blocks from CSR arrays: offsets[n_blocks+1], markers[offsets[n_blocks]] */
EXPORT void* refb_blocks_from_csr(unsigned n_blocks, const unsigned* offsets, const unsigned* markers) {
    gsc_MarkerBlocks* b = malloc(sizeof *b);
    b->num_blocks = n_blocks;
    b->num_markers_in_block = malloc(sizeof(GSC_GENOLEN_T) * (n_blocks ? n_blocks : 1));
    b->markers_in_block = malloc(sizeof(GSC_GENOLEN_T*) * (n_blocks ? n_blocks : 1));
    for (unsigned i = 0; i < n_blocks; ++i) {
        unsigned len = offsets[i + 1] - offsets[i];
        b->num_markers_in_block[i] = len;
        b->markers_in_block[i] = len ? malloc(sizeof(GSC_GENOLEN_T) * len) : NULL;
        for (unsigned k = 0; k < len; ++k) b->markers_in_block[i][k] = markers[offsets[i] + k];
    }
    return b;
}
EXPORT unsigned refb_blocks_n(void* b) { return ((gsc_MarkerBlocks*) b)->num_blocks; }
EXPORT unsigned refb_blocks_total(void* bv) {
    gsc_MarkerBlocks* b = (gsc_MarkerBlocks*) bv; unsigned t = 0;
    for (unsigned i = 0; i < b->num_blocks; ++i) t += b->num_markers_in_block[i];
    return t;
}
EXPORT void refb_blocks_export(void* bv, unsigned* offsets, unsigned* markers) {
    gsc_MarkerBlocks* b = (gsc_MarkerBlocks*) bv; unsigned t = 0;
    for (unsigned i = 0; i < b->num_blocks; ++i) {
        offsets[i] = t;
        for (unsigned k = 0; k < b->num_markers_in_block[i]; ++k) markers[t++] = b->markers_in_block[i][k];
    }
    offsets[b->num_blocks] = t;
}
EXPORT void refb_blocks_free(void* b) { gsc_delete_markerblocks((gsc_MarkerBlocks*) b); free(b); }

/* out is row-major (2*n_genotypes) x n_blocks; returns number of rows */
EXPORT unsigned refb_calculate_local_bvs(void* d, unsigned group, void* blocks, unsigned eff_id,
                                         unsigned cap_rows, double* out) {
    gsc_MarkerBlocks* b = (gsc_MarkerBlocks*) blocks;
    gsc_DecimalMatrix m = gsc_calculate_local_bvs((gsc_SimData*) d, (gsc_GroupNum){.num = group}, *b,
                                                  (gsc_EffectID){.id = eff_id});
    unsigned rows = m.dim1;
    if (out) for (unsigned i = 0; i < rows && i < cap_rows; ++i)
        memcpy(out + (size_t) i * m.dim2, m.matrix[i], sizeof(double) * m.dim2);
    gsc_delete_dmatrix(&m);
    return rows;
}

/* out is row-major n_genotypes x n_markers; returns number of rows */
EXPORT unsigned refb_calculate_allele_counts(void* d, unsigned group, char allele, unsigned cap_rows, double* out) {
    gsc_DecimalMatrix m = gsc_calculate_allele_counts((gsc_SimData*) d, (gsc_GroupNum){.num = group}, allele);
    unsigned rows = m.dim1;
    if (out) for (unsigned i = 0; i < rows && i < cap_rows; ++i)
        memcpy(out + (size_t) i * m.dim2, m.matrix[i], sizeof(double) * m.dim2);
    gsc_delete_dmatrix(&m);
    return rows;
}

EXPORT unsigned refb_split_by_bv(void* d, unsigned group, unsigned eff_id, unsigned top_n, int low_is_best) {
    return gsc_split_by_bv((gsc_SimData*) d, (gsc_GroupNum){.num = group}, (gsc_EffectID){.id = eff_id},
                           top_n, low_is_best != 0).num;
}
EXPORT int refb_recombinations_from_file(void* d, const char* input_file, const char* output_file, int window_len, int certain) {
    return gsc_calculate_recombinations_from_file((gsc_SimData*) d, input_file, output_file, window_len, certain);
}
/* group-split utilities of the reference (core.c:2980-3784); kind: 0 families, 1 / 2 half-sibs on parent 1 / 2, 3 individuals,
 * 4 evenly into two, 5 evenly into n, 6 buckets (counts), 7 randomly into two, 8 randomly into n, 9 by probabilities */
EXPORT unsigned refb_split(void* dv, int kind, unsigned group, unsigned n, const unsigned* counts, const double* probs, unsigned cap, unsigned* out) {
    gsc_SimData* d = (gsc_SimData*) dv;
    const gsc_GroupNum g = { .num = group };
    gsc_GroupNum res[cap ? cap : 1];
    unsigned made = 0;
    for (unsigned i = 0; i < cap; ++i) res[i].num = 0;
    switch (kind) {
        case 0: made = gsc_split_into_families(d, g, cap, res); break;
        case 1: case 2: made = gsc_split_into_halfsib_families(d, g, kind, cap, res); break;
        case 3: made = gsc_split_into_individuals(d, g, cap, res); break;
        case 4: res[0] = gsc_split_evenly_into_two(d, g); made = res[0].num != 0; break;
        case 5: made = gsc_split_evenly_into_n(d, g, n, res); break;
        case 6: made = gsc_split_into_buckets(d, g, n, counts, res); break;
        case 7: res[0] = gsc_split_randomly_into_two(d, g); made = res[0].num != 0; break;
        case 8: made = gsc_split_randomly_into_n(d, g, n, res); break;
        case 9: made = gsc_split_by_probabilities(d, g, n, probs, res); break;
        default: break;
    }
    for (unsigned i = 0; i < cap; ++i) out[i] = res[i].num;
    return made;
}
EXPORT unsigned refb_make_group_from(void* d, unsigned n, const unsigned* indexes) {
    return gsc_make_group_from((gsc_SimData*) d, n, indexes).num;
}
EXPORT unsigned refb_combine_groups(void* d, unsigned n, const unsigned* groups) {
    gsc_GroupNum g[n];
    for (unsigned i = 0; i < n; ++i) g[i].num = groups[i];
    return gsc_combine_groups((gsc_SimData*) d, n, g).num;
}
EXPORT void refb_delete_group(void* d, unsigned group) { gsc_delete_group((gsc_SimData*) d, (gsc_GroupNum){.num = group}); }

/* ---------------- savers (text boundary) ---------------- */

EXPORT void refb_save_genotypes(void* d, const char* f, unsigned group, int markers_as_rows) {
    gsc_save_genotypes(f, (gsc_SimData*) d, (gsc_GroupNum){.num = group}, markers_as_rows != 0);
}
EXPORT void refb_save_bvs(void* d, const char* f, unsigned group, unsigned eff_id) {
    gsc_save_bvs(f, (gsc_SimData*) d, (gsc_GroupNum){.num = group}, (gsc_EffectID){.id = eff_id});
}
EXPORT void refb_save_allele_counts(void* d, const char* f, unsigned group, char allele, int markers_as_rows) {
    gsc_save_allele_counts(f, (gsc_SimData*) d, (gsc_GroupNum){.num = group}, allele, markers_as_rows != 0);
}
EXPORT void refb_save_local_bvs(void* d, const char* f, unsigned group, void* blocks, unsigned eff_id, int headers) {
    gsc_save_local_bvs(f, (gsc_SimData*) d, (gsc_GroupNum){.num = group}, *(gsc_MarkerBlocks*) blocks,
                       (gsc_EffectID){.id = eff_id}, headers != 0);
}
EXPORT void refb_save_pedigrees(void* d, const char* f, unsigned group, int full) {
    gsc_save_pedigrees(f, (gsc_SimData*) d, (gsc_GroupNum){.num = group}, full != 0);
}
