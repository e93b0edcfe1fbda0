/* gsb200_compat.h — the reference's own names and struct types for the hot path, over the
 * device-backed mirror (include/gsb200_gsc.h).
 *
 * `r-wrappers.c` calls the core through `gsc_*` functions that take and return the reference's
 * struct types — `gsc_GroupNum{.num}` (core.h:561-563), `gsc_MapID` / `gsc_EffectID`,
 * `gsc_GenOptions` (core.h:629-669), `gsc_DecimalMatrix{double** matrix, dim1, dim2}`
 * (core.h:536-540), ragged `gsc_MarkerBlocks` (core.h:515-528) — and reads a few `gsc_SimData`
 * fields directly (`n_eff_sets`, `eff_set_ids`, `genome.n_maps`, `genome.map_ids`,
 * `genome.n_markers`, `m`; r-wrappers.c:44-88, 1073-1104, 1996-2020).  This header declares exactly
 * those, with the same names, member names and argument orders, so that the hot-path `SXP_*` entry
 * points of an UNMODIFIED r-wrappers.c compile against it: oracle/Makefile `wrappers` extracts them
 * from the reference source where it lies and builds them against this header and a minimal
 * Rinternals (tests/test_gpu_wrappers.py calls them).  INTEGRATION.md lists what was compiled.
 *
 * The short names (`SimData`, `make_random_crosses`, ...) follow core.h:186-420; define
 * GSC_NO_SHORT_NAMES to leave them out, as in the reference.
 */
#ifndef GSB200_COMPAT_H
#define GSB200_COMPAT_H

#include "gsb200_gsc.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#ifndef GSC_ID_T
#define GSC_ID_T unsigned int          /* core.h:59-62 */
#define GSC_LOCALX_T unsigned int
#define GSC_GLOBALX_T unsigned int
#define GSC_GENOLEN_T unsigned int
#endif
#define GSC_NA_ID 0                    /* core.h:108-169 */
#define GSC_NA_GLOBALX ((GSC_GLOBALX_T) -1)
#define GSC_NA (-1)
#define GSC_TRUE 1
#define GSC_FALSE 0

typedef struct { GSC_ID_T id; } gsc_PedigreeID;      /* core.h:548-550 */
typedef struct { GSC_ID_T num; } gsc_GroupNum;       /* core.h:561-563 */
typedef struct { GSC_ID_T id; } gsc_EffectID;        /* core.h:575-577 */
typedef struct { GSC_ID_T id; } gsc_LabelID;
typedef struct { GSC_ID_T id; } gsc_MapID;
#define GSC_NO_PEDIGREE  (gsc_PedigreeID){.id = GSC_NA_ID}
#define GSC_NO_GROUP     (gsc_GroupNum){.num = GSC_NA_ID}
#define GSC_NO_EFFECTSET (gsc_EffectID){.id = GSC_NA_ID}
#define GSC_NO_MAP       (gsc_MapID){.id = GSC_NA_ID}

/* core.h:515-528 */
typedef struct {
    GSC_GENOLEN_T num_blocks;
    GSC_GENOLEN_T* num_markers_in_block;
    GSC_GENOLEN_T** markers_in_block;
} gsc_MarkerBlocks;

/* core.h:536-540: every row separately allocated, as gsc_generate_zero_dmatrix does (core.c:419-429) */
typedef struct {
    double** matrix;
    unsigned int dim1;
    unsigned int dim2;
} gsc_DecimalMatrix;

/* core.h:629-669 */
typedef struct {
    _Bool will_name_offspring;
    const char* offspring_name_prefix;
    GSC_GLOBALX_T family_size;
    _Bool will_track_pedigree;
    _Bool will_allocate_ids;
    const char* filename_prefix;
    _Bool will_save_pedigree_to_file;
    gsc_EffectID will_save_bvs_to_file;
    _Bool will_save_alleles_to_file;
    _Bool will_save_to_simdata;
} gsc_GenOptions;
extern const gsc_GenOptions GSC_BASIC_OPT;          /* core.c:11-23 */

typedef struct gsc_SimData gsc_SimData;
/* Stands where the reference's linked list of genotype blocks is handed around (d->m): the payload
 * lives on the device, so this is only a handle back to the simulation. */
typedef struct gsc_AlleleMatrix { gsc_SimData* owner; } gsc_AlleleMatrix;

/* The members of gsc_SimData (core.h:863-901) that r-wrappers.c reads on the hot path, kept current
 * by gsc_compat_refresh; `engine` is the device-backed simulation. */
struct gsc_SimData {
    GSC_ID_T n_eff_sets;
    gsc_EffectID* eff_set_ids;
    struct {
        GSC_GENOLEN_T n_markers;
        char** marker_names;
        GSC_ID_T n_maps;
        gsc_MapID* map_ids;
    } genome;
    gsc_AlleleMatrix* m;
    gsc_PedigreeID current_id;
    gsb_SimData* engine;
    gsc_AlleleMatrix m_handle;
};

/* a facade over an existing device-backed simulation (owned by the caller) */
gsc_SimData* gsc_compat_attach(gsb_SimData* engine);
void gsc_compat_refresh(gsc_SimData* d);             /* after maps / effect sets / genotypes changed */
void gsc_compat_detach(gsc_SimData* d);

/* ---- the hot-path API under its reference names (core.h:1636-1797) */
GSC_GLOBALX_T gsc_get_group_size(const gsc_SimData* d, const gsc_GroupNum group_id);
GSC_GLOBALX_T gsc_get_group_bvs(const gsc_SimData* d, const gsc_GroupNum group_id, const gsc_EffectID effID, GSC_GLOBALX_T group_size, double* output);
GSC_GLOBALX_T gsc_get_index_of_name(const gsc_AlleleMatrix* start, const char* name);
void gsc_delete_group(gsc_SimData* d, const gsc_GroupNum group_id);

gsc_GroupNum gsc_make_random_crosses(gsc_SimData* d, const gsc_GroupNum from_group, const GSC_GLOBALX_T n_crosses,
                                     const GSC_GLOBALX_T cap, const gsc_MapID which_map, const gsc_GenOptions g);
gsc_GroupNum gsc_make_random_crosses_between(gsc_SimData* d, const gsc_GroupNum group1, const gsc_GroupNum group2,
                                             const GSC_GLOBALX_T n_crosses, const GSC_GLOBALX_T cap1, const GSC_GLOBALX_T cap2,
                                             const gsc_MapID map1, const gsc_MapID map2, const gsc_GenOptions g);
gsc_GroupNum gsc_make_targeted_crosses(gsc_SimData* d, const GSC_GLOBALX_T n_combinations, const GSC_GLOBALX_T* firstParents,
                                       const GSC_GLOBALX_T* secondParents, const gsc_MapID map1, const gsc_MapID map2, const gsc_GenOptions g);
gsc_GroupNum gsc_self_n_times(gsc_SimData* d, const unsigned int n, const gsc_GroupNum group, const gsc_MapID which_map, const gsc_GenOptions g);
gsc_GroupNum gsc_make_doubled_haploids(gsc_SimData* d, const gsc_GroupNum group, const gsc_MapID which_map, const gsc_GenOptions g);
gsc_GroupNum gsc_make_clones(gsc_SimData* d, const gsc_GroupNum group, const _Bool inherit_names, const gsc_GenOptions g);
gsc_GroupNum gsc_make_all_unidirectional_crosses(gsc_SimData* d, const gsc_GroupNum from_group, const gsc_MapID mapID, const gsc_GenOptions g);

gsc_GroupNum gsc_split_by_bv(gsc_SimData* d, const gsc_GroupNum group, const gsc_EffectID effID, const GSC_GLOBALX_T top_n, const _Bool lowIsBest);
int gsc_calculate_recombinations_from_file(gsc_SimData* d, const char* input_file, const char* output_file, int window_len, int certain);   /* core.h:1889 */
/* group-split utilities, core.h:1469-1520 */
unsigned int gsc_split_into_families(gsc_SimData* d, const gsc_GroupNum group_id, unsigned int maxentries_results, gsc_GroupNum* results);
unsigned int gsc_split_into_halfsib_families(gsc_SimData* d, const gsc_GroupNum group_id, const int parent, unsigned int maxentries_results, gsc_GroupNum* results);
unsigned int gsc_split_into_individuals(gsc_SimData* d, const gsc_GroupNum group_id, unsigned int maxentries_results, gsc_GroupNum* results);
gsc_GroupNum gsc_split_evenly_into_two(gsc_SimData* d, const gsc_GroupNum group_id);
unsigned int gsc_split_evenly_into_n(gsc_SimData* d, const gsc_GroupNum group_id, const unsigned int n, gsc_GroupNum* results);
unsigned int gsc_split_into_buckets(gsc_SimData* d, const gsc_GroupNum group_id, const unsigned int n, const GSC_GLOBALX_T* counts, gsc_GroupNum* results);
gsc_GroupNum gsc_split_randomly_into_two(gsc_SimData* d, const gsc_GroupNum group_id);
unsigned int gsc_split_randomly_into_n(gsc_SimData* d, const gsc_GroupNum group_id, const unsigned int n, gsc_GroupNum* results);
unsigned int gsc_split_by_probabilities(gsc_SimData* d, const gsc_GroupNum group_id, const unsigned int n, const double* probs, gsc_GroupNum* results);
gsc_DecimalMatrix gsc_calculate_bvs(const gsc_SimData* d, const gsc_GroupNum group, const gsc_EffectID effID);
gsc_DecimalMatrix gsc_calculate_allele_counts(const gsc_SimData* d, const gsc_GroupNum group, const char allele);
gsc_DecimalMatrix gsc_calculate_local_bvs(const gsc_SimData* d, const gsc_GroupNum group, const gsc_MarkerBlocks b, const gsc_EffectID effID);
gsc_MarkerBlocks gsc_create_evenlength_blocks_each_chr(const gsc_SimData* d, const gsc_MapID mapid, const GSC_ID_T n);
void gsc_delete_dmatrix(gsc_DecimalMatrix* m);
void gsc_delete_markerblocks(gsc_MarkerBlocks* b);

void gsc_save_genotypes(const char* fname, const gsc_SimData* d, const gsc_GroupNum groupID, const _Bool markers_as_rows);
void gsc_save_allele_counts(const char* fname, const gsc_SimData* d, const gsc_GroupNum groupID, const char allele, const _Bool markers_as_rows);
void gsc_save_pedigrees(const char* fname, const gsc_SimData* d, const gsc_GroupNum groupID, const _Bool full_pedigree);
void gsc_save_bvs(const char* fname, const gsc_SimData* d, const gsc_GroupNum groupID, const gsc_EffectID effID);
void gsc_save_local_bvs(const char* fname, const gsc_SimData* d, const gsc_GroupNum groupID, const gsc_MarkerBlocks b,
                        const gsc_EffectID effID, const _Bool headers);

#ifndef GSC_NO_SHORT_NAMES                           /* core.h:186-420 */
typedef gsc_SimData SimData;
typedef gsc_AlleleMatrix AlleleMatrix;
typedef gsc_GenOptions GenOptions;
typedef gsc_MarkerBlocks MarkerBlocks;
typedef gsc_DecimalMatrix DecimalMatrix;
typedef gsc_PedigreeID PedigreeID;
typedef gsc_GroupNum GroupNum;
typedef gsc_EffectID EffectID;
typedef gsc_LabelID LabelID;
typedef gsc_MapID MapID;
#define NO_PEDIGREE GSC_NO_PEDIGREE
#define NO_GROUP GSC_NO_GROUP
#define NO_EFFECTSET GSC_NO_EFFECTSET
#define NO_MAP GSC_NO_MAP
#define BASIC_OPT GSC_BASIC_OPT
#define get_group_size gsc_get_group_size
#define get_group_bvs gsc_get_group_bvs
#define get_index_of_name gsc_get_index_of_name
#define delete_group gsc_delete_group
#define make_random_crosses gsc_make_random_crosses
#define make_random_crosses_between gsc_make_random_crosses_between
#define make_targeted_crosses gsc_make_targeted_crosses
#define self_n_times gsc_self_n_times
#define make_doubled_haploids gsc_make_doubled_haploids
#define make_clones gsc_make_clones
#define make_all_unidirectional_crosses gsc_make_all_unidirectional_crosses
#define split_by_bv gsc_split_by_bv
#define calculate_recombinations_from_file gsc_calculate_recombinations_from_file
#define split_into_families gsc_split_into_families
#define split_into_halfsib_families gsc_split_into_halfsib_families
#define split_into_individuals gsc_split_into_individuals
#define split_evenly_into_two gsc_split_evenly_into_two
#define split_evenly_into_n gsc_split_evenly_into_n
#define split_into_buckets gsc_split_into_buckets
#define split_randomly_into_two gsc_split_randomly_into_two
#define split_randomly_into_n gsc_split_randomly_into_n
#define split_by_probabilities gsc_split_by_probabilities
#define calculate_bvs gsc_calculate_bvs
#define calculate_allele_counts gsc_calculate_allele_counts
#define calculate_local_bvs gsc_calculate_local_bvs
#define create_evenlength_blocks_each_chr gsc_create_evenlength_blocks_each_chr
#define delete_dmatrix gsc_delete_dmatrix
#define delete_markerblocks gsc_delete_markerblocks
#define save_genotypes gsc_save_genotypes
#define save_allele_counts gsc_save_allele_counts
#define save_pedigrees gsc_save_pedigrees
#define save_bvs gsc_save_bvs
#define save_local_bvs gsc_save_local_bvs
#endif

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
