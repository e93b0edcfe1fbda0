/* gsb_compat.c — include/gsb200_compat.h: the reference's names and struct types over the mirror.
 * Pure adaptation: struct-wrapped ids <-> plain ids, gsc_DecimalMatrix's row pointers <-> the flat
 * matrix, ragged gsc_MarkerBlocks <-> CSR.  No arithmetic lives here. */
#define _GNU_SOURCE
#include "../../include/gsb200_compat.h"

#include <stdlib.h>
#include <string.h>

const gsc_GenOptions GSC_BASIC_OPT = {
    .will_name_offspring = 0, .offspring_name_prefix = NULL, .family_size = 1,
    .will_track_pedigree = 0, .will_allocate_ids = 1, .filename_prefix = NULL,
    .will_save_pedigree_to_file = 0, .will_save_bvs_to_file = { 0 }, .will_save_alleles_to_file = 0,
    .will_save_to_simdata = 1
};

static gsb_GenOptions opts(const gsc_GenOptions g) {
    gsb_GenOptions o = { g.will_name_offspring, g.offspring_name_prefix, g.family_size, g.will_track_pedigree,
                         g.will_allocate_ids, g.filename_prefix, g.will_save_pedigree_to_file, g.will_save_bvs_to_file.id,
                         g.will_save_alleles_to_file, g.will_save_to_simdata };
    return o;
}

gsc_SimData* gsc_compat_attach(gsb_SimData* engine) {
    if (!engine) return NULL;
    gsc_SimData* d = calloc(1, sizeof *d);
    d->engine = engine;
    d->m_handle.owner = d;
    d->m = &d->m_handle;
    gsc_compat_refresh(d);
    return d;
}

void gsc_compat_refresh(gsc_SimData* d) {
    free(d->eff_set_ids);
    free(d->genome.map_ids);
    unsigned n = gsb_get_eff_set_ids(d->engine, 0, NULL);
    d->eff_set_ids = calloc(n ? n : 1, sizeof *d->eff_set_ids);
    d->n_eff_sets = gsb_get_eff_set_ids(d->engine, n, (gsb_EffectID*) d->eff_set_ids);
    n = gsb_get_map_ids(d->engine, 0, NULL);
    d->genome.map_ids = calloc(n ? n : 1, sizeof *d->genome.map_ids);
    d->genome.n_maps = gsb_get_map_ids(d->engine, n, (gsb_MapID*) d->genome.map_ids);
    d->genome.n_markers = gsb_get_n_markers(d->engine);
    d->genome.marker_names = (char**) gsb_get_marker_names(d->engine);
    d->current_id.id = gsb_get_current_id(d->engine);
}

void gsc_compat_detach(gsc_SimData* d) {
    if (!d) return;
    free(d->eff_set_ids);
    free(d->genome.map_ids);
    free(d);
}

static gsc_GroupNum made(gsc_SimData* d, gsb_GroupNum g) {
    d->current_id.id = gsb_get_current_id(d->engine);
    return (gsc_GroupNum){ .num = g };
}

GSC_GLOBALX_T gsc_get_group_size(const gsc_SimData* d, const gsc_GroupNum g) { return gsb_get_group_size(d->engine, g.num); }
GSC_GLOBALX_T gsc_get_group_bvs(const gsc_SimData* d, const gsc_GroupNum g, const gsc_EffectID e, GSC_GLOBALX_T n, double* out) {
    return gsb_get_group_bvs(d->engine, g.num, e.id, n, out);
}
GSC_GLOBALX_T gsc_get_index_of_name(const gsc_AlleleMatrix* start, const char* name) { return gsb_get_index_of_name(start->owner->engine, name); }
void gsc_delete_group(gsc_SimData* d, const gsc_GroupNum g) { gsb_delete_group(d->engine, g.num); }

gsc_GroupNum gsc_make_random_crosses(gsc_SimData* d, const gsc_GroupNum from, const GSC_GLOBALX_T n, const GSC_GLOBALX_T cap,
                                     const gsc_MapID map, const gsc_GenOptions g) {
    return made(d, gsb_make_random_crosses(d->engine, from.num, n, cap, map.id, opts(g)));
}
gsc_GroupNum gsc_make_random_crosses_between(gsc_SimData* d, const gsc_GroupNum g1, const gsc_GroupNum g2, const GSC_GLOBALX_T n,
                                             const GSC_GLOBALX_T cap1, const GSC_GLOBALX_T cap2, const gsc_MapID map1, const gsc_MapID map2,
                                             const gsc_GenOptions g) {
    return made(d, gsb_make_random_crosses_between(d->engine, g1.num, g2.num, n, cap1, cap2, map1.id, map2.id, opts(g)));
}
gsc_GroupNum gsc_make_targeted_crosses(gsc_SimData* d, const GSC_GLOBALX_T n, const GSC_GLOBALX_T* first, const GSC_GLOBALX_T* second,
                                       const gsc_MapID map1, const gsc_MapID map2, const gsc_GenOptions g) {
    return made(d, gsb_make_targeted_crosses(d->engine, n, first, second, map1.id, map2.id, opts(g)));
}
gsc_GroupNum gsc_self_n_times(gsc_SimData* d, const unsigned int n, const gsc_GroupNum group, const gsc_MapID map, const gsc_GenOptions g) {
    return made(d, gsb_self_n_times(d->engine, n, group.num, map.id, opts(g)));
}
gsc_GroupNum gsc_make_doubled_haploids(gsc_SimData* d, const gsc_GroupNum group, const gsc_MapID map, const gsc_GenOptions g) {
    return made(d, gsb_make_doubled_haploids(d->engine, group.num, map.id, opts(g)));
}
gsc_GroupNum gsc_make_clones(gsc_SimData* d, const gsc_GroupNum group, const _Bool inherit_names, const gsc_GenOptions g) {
    return made(d, gsb_make_clones(d->engine, group.num, inherit_names, opts(g)));
}
gsc_GroupNum gsc_make_all_unidirectional_crosses(gsc_SimData* d, const gsc_GroupNum from, const gsc_MapID map, const gsc_GenOptions g) {
    return made(d, gsb_make_all_unidirectional_crosses(d->engine, from.num, map.id, opts(g)));
}
gsc_GroupNum gsc_split_by_bv(gsc_SimData* d, const gsc_GroupNum group, const gsc_EffectID e, const GSC_GLOBALX_T top_n, const _Bool low) {
    return (gsc_GroupNum){ .num = gsb_split_by_bv(d->engine, group.num, e.id, top_n, low) };
}

int gsc_calculate_recombinations_from_file(gsc_SimData* d, const char* input_file, const char* output_file, int window_len, int certain) {
    return gsb_calculate_recombinations_from_file(d->engine, input_file, output_file, window_len, certain);
}

/* gsc_GroupNum is a struct of one unsigned (core.h:561-563): an array of them is an array of unsigned */
_Static_assert(sizeof(gsc_GroupNum) == sizeof(unsigned), "gsc_GroupNum arrays are passed through as unsigned arrays");
unsigned int gsc_split_into_families(gsc_SimData* d, const gsc_GroupNum g, unsigned int cap, gsc_GroupNum* results) {
    return gsb_split_into_families(d->engine, g.num, cap, (unsigned*) results);
}
unsigned int gsc_split_into_halfsib_families(gsc_SimData* d, const gsc_GroupNum g, const int parent, unsigned int cap, gsc_GroupNum* results) {
    return gsb_split_into_halfsib_families(d->engine, g.num, parent, cap, (unsigned*) results);
}
unsigned int gsc_split_into_individuals(gsc_SimData* d, const gsc_GroupNum g, unsigned int cap, gsc_GroupNum* results) {
    return gsb_split_into_individuals(d->engine, g.num, cap, (unsigned*) results);
}
gsc_GroupNum gsc_split_evenly_into_two(gsc_SimData* d, const gsc_GroupNum g) { return (gsc_GroupNum){ .num = gsb_split_evenly_into_two(d->engine, g.num) }; }
unsigned int gsc_split_evenly_into_n(gsc_SimData* d, const gsc_GroupNum g, const unsigned int n, gsc_GroupNum* results) {
    return gsb_split_evenly_into_n(d->engine, g.num, n, (unsigned*) results);
}
unsigned int gsc_split_into_buckets(gsc_SimData* d, const gsc_GroupNum g, const unsigned int n, const GSC_GLOBALX_T* counts, gsc_GroupNum* results) {
    return gsb_split_into_buckets(d->engine, g.num, n, counts, (unsigned*) results);
}
gsc_GroupNum gsc_split_randomly_into_two(gsc_SimData* d, const gsc_GroupNum g) { return (gsc_GroupNum){ .num = gsb_split_randomly_into_two(d->engine, g.num) }; }
unsigned int gsc_split_randomly_into_n(gsc_SimData* d, const gsc_GroupNum g, const unsigned int n, gsc_GroupNum* results) {
    return gsb_split_randomly_into_n(d->engine, g.num, n, (unsigned*) results);
}
unsigned int gsc_split_by_probabilities(gsc_SimData* d, const gsc_GroupNum g, const unsigned int n, const double* probs, gsc_GroupNum* results) {
    return gsb_split_by_probabilities(d->engine, g.num, n, probs, (unsigned*) results);
}

/* flat row-major -> one allocation per row, as the reference hands matrices out (core.c:419-429) */
static gsc_DecimalMatrix rows_of(gsb_DecimalMatrix f) {
    gsc_DecimalMatrix m = { NULL, 0, 0 };
    if (!f.data || !f.dim1 || !f.dim2) { gsb_delete_dmatrix(&f); return m; }
    m.dim1 = f.dim1; m.dim2 = f.dim2;
    m.matrix = malloc(sizeof(double*) * f.dim1);
    for (unsigned i = 0; i < f.dim1; ++i) {
        m.matrix[i] = malloc(sizeof(double) * f.dim2);
        memcpy(m.matrix[i], f.data + (size_t) i * f.dim2, sizeof(double) * f.dim2);
    }
    gsb_delete_dmatrix(&f);
    return m;
}

void gsc_delete_dmatrix(gsc_DecimalMatrix* m) {              /* core.c:4628-4641 */
    if (m->matrix) {
        for (unsigned i = 0; i < m->dim1; ++i) free(m->matrix[i]);
        free(m->matrix);
        m->matrix = NULL;
    }
    m->dim1 = m->dim2 = 0;
}

static gsb_MarkerBlocks csr_of(const gsc_MarkerBlocks b) {
    gsb_MarkerBlocks c = { b.num_blocks, calloc((size_t) b.num_blocks + 1, sizeof(unsigned)), NULL };
    for (unsigned i = 0; i < b.num_blocks; ++i) c.offsets[i + 1] = c.offsets[i] + b.num_markers_in_block[i];
    c.markers = malloc(sizeof(unsigned) * (c.offsets[b.num_blocks] ? c.offsets[b.num_blocks] : 1));
    for (unsigned i = 0; i < b.num_blocks; ++i)
        memcpy(c.markers + c.offsets[i], b.markers_in_block[i], sizeof(unsigned) * b.num_markers_in_block[i]);
    return c;
}

gsc_DecimalMatrix gsc_calculate_bvs(const gsc_SimData* d, const gsc_GroupNum group, const gsc_EffectID e) {
    return rows_of(gsb_calculate_bvs(d->engine, group.num, e.id));
}
gsc_DecimalMatrix gsc_calculate_allele_counts(const gsc_SimData* d, const gsc_GroupNum group, const char allele) {
    return rows_of(gsb_calculate_allele_counts(d->engine, group.num, allele));
}
gsc_DecimalMatrix gsc_calculate_local_bvs(const gsc_SimData* d, const gsc_GroupNum group, const gsc_MarkerBlocks b, const gsc_EffectID e) {
    gsb_MarkerBlocks c = csr_of(b);
    gsc_DecimalMatrix m = rows_of(gsb_calculate_local_bvs(d->engine, group.num, c, e.id));
    gsb_delete_markerblocks(&c);
    return m;
}

gsc_MarkerBlocks gsc_create_evenlength_blocks_each_chr(const gsc_SimData* d, const gsc_MapID mapid, const GSC_ID_T n) {
    gsb_MarkerBlocks c = gsb_create_evenlength_blocks_each_chr(d->engine, mapid.id, n);
    gsc_MarkerBlocks b = { c.num_blocks, NULL, NULL };
    if (c.num_blocks) {
        b.num_markers_in_block = malloc(sizeof(GSC_GENOLEN_T) * c.num_blocks);
        b.markers_in_block = malloc(sizeof(GSC_GENOLEN_T*) * c.num_blocks);
        for (unsigned i = 0; i < c.num_blocks; ++i) {
            const unsigned len = c.offsets[i + 1] - c.offsets[i];
            b.num_markers_in_block[i] = len;
            b.markers_in_block[i] = len ? malloc(sizeof(GSC_GENOLEN_T) * len) : NULL;
            if (len) memcpy(b.markers_in_block[i], c.markers + c.offsets[i], sizeof(GSC_GENOLEN_T) * len);
        }
    }
    gsb_delete_markerblocks(&c);
    return b;
}

void gsc_delete_markerblocks(gsc_MarkerBlocks* b) {          /* core.c:4757-4766 */
    for (unsigned i = 0; i < b->num_blocks; ++i) free(b->markers_in_block[i]);
    free(b->markers_in_block);
    free(b->num_markers_in_block);
    b->markers_in_block = NULL; b->num_markers_in_block = NULL; b->num_blocks = 0;
}

void gsc_save_genotypes(const char* f, const gsc_SimData* d, const gsc_GroupNum g, const _Bool markers_as_rows) {
    gsb_save_genotypes(f, d->engine, g.num, markers_as_rows);
}
void gsc_save_allele_counts(const char* f, const gsc_SimData* d, const gsc_GroupNum g, const char allele, const _Bool markers_as_rows) {
    gsb_save_allele_counts(f, d->engine, g.num, allele, markers_as_rows);
}
void gsc_save_pedigrees(const char* f, const gsc_SimData* d, const gsc_GroupNum g, const _Bool full) {
    gsb_save_pedigrees(f, d->engine, g.num, full);
}
void gsc_save_bvs(const char* f, const gsc_SimData* d, const gsc_GroupNum g, const gsc_EffectID e) { gsb_save_bvs(f, d->engine, g.num, e.id); }
void gsc_save_local_bvs(const char* f, const gsc_SimData* d, const gsc_GroupNum g, const gsc_MarkerBlocks b, const gsc_EffectID e, const _Bool headers) {
    gsb_MarkerBlocks c = csr_of(b);
    gsb_save_local_bvs(f, d->engine, g.num, c, e.id, headers);
    gsb_delete_markerblocks(&c);
}
