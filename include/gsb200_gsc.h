/* gsb200_gsc.h — host-side mirror of genomicSimulation's C API for the hot path.
 *
 * Plain C over include/gsb200.h.  Each function keeps the name (gsc_ -> gsb_), argument order,
 * argument meaning and soft-error behaviour (a "NOTE! ..." message and a 0 / empty return) of
 * the reference function it stands in for in `src/sim-operations.c` ("core.c") /
 * `src/sim-operations.h` ("core.h"), so that the r-wrappers.c entry points (SXP_make_random_crosses,
 * SXP_see_group_data, ...) can call these instead.  What differs is where the
 * genotype payload lives: gsc_AlleleMatrix's char[2M] per genotype (core.h:776-806) becomes a
 * row of the device-resident bit-packed store; everything that is metadata (ids, pedigrees,
 * group numbers, names, storage order = "global index") stays on the host, in the
 * reference's order.
 *
 * Random draws.  The reference pulls every draw from R's stream in one interleaved order
 * (core.c:8299-8332).  Two modes:
 *   GSB_DRAWS_HOST    every draw — parent picks AND crossover counts/positions/strands — is
 *                     pulled from the host draw source (gsb_set_draw_source: R's unif_rand /
 *                     Rf_rpois in the package; a replay tape in the parity tests) in exactly
 *                     the reference's order, then shipped to the GPU as a tape.  Results are
 *                     bit-identical to the reference's for the same stream.
 *   GSB_DRAWS_NATIVE  meiosis draws come from the device Philox stream keyed by (seed, call
 *                     number) and indexed by offspring.  Random crosses without a usage cap
 *                     draw their parent pairs from the same stream on the device (same pick
 *                     rule as core.c:8556-8581) — only the candidate list goes up and the
 *                     picked pairs (for pedigrees) come down; capped random crosses pick on
 *                     the host from the host source, in the reference's order.
 */
#ifndef GSB200_GSC_H
#define GSB200_GSC_H

#include <stddef.h>
#include <stdint.h>
#include "gsb200.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef unsigned int gsb_GroupNum;   /* gsc_GroupNum.num  (core.h:561-563); 0 = GSC_NO_GROUP */
typedef unsigned int gsb_MapID;      /* gsc_MapID.id; 0 = GSC_NO_MAP -> first map (core.c:8514-8520) */
typedef unsigned int gsb_EffectID;   /* gsc_EffectID.id; 0 = GSC_NO_EFFECTSET */
#define GSB_NA_GLOBALX ((unsigned int) -1)   /* GSC_NA_GLOBALX (core.h:160) */

/* gsc_GenOptions (core.h:629-669), same fields in the same order. */
typedef struct {
    _Bool will_name_offspring;
    const char* offspring_name_prefix;
    unsigned int family_size;
    _Bool will_track_pedigree;
    _Bool will_allocate_ids;
    const char* filename_prefix;
    _Bool will_save_pedigree_to_file;
    gsb_EffectID will_save_bvs_to_file;
    _Bool will_save_alleles_to_file;
    _Bool will_save_to_simdata;
} gsb_GenOptions;
extern const gsb_GenOptions GSB_BASIC_OPT;   /* GSC_BASIC_OPT (core.c:11-23) */

/* gsc_DecimalMatrix (core.h:536-540) with flat row-major storage. Free with gsb_delete_dmatrix. */
typedef struct {
    unsigned int dim1, dim2;
    double* data;
} gsb_DecimalMatrix;

/* gsc_MarkerBlocks (core.h:515-528) as CSR: block b = markers[offsets[b] .. offsets[b+1]). */
typedef struct {
    unsigned int num_blocks;
    unsigned int* offsets;
    unsigned int* markers;
} gsb_MarkerBlocks;

typedef struct gsb_SimData gsb_SimData;

enum { GSB_DRAWS_HOST = 0, GSB_DRAWS_NATIVE = 1 };

/* ---- lifetime (gsc_create_empty_simdata core.c:113-132, gsc_delete_simdata core.c:5047) */
gsb_SimData* gsb_create_simdata(int device_ordinal, unsigned int n_markers);
void gsb_delete_simdata(gsb_SimData* d);
gsb200_ctx* gsb_context(gsb_SimData* d);
/* status of the last call into the CUDA engine (GSB200_OK, ...) and its message */
int gsb_last_status(const gsb_SimData* d);
const char* gsb_last_message(const gsb_SimData* d);
/* where "NOTE! ..." messages go (Rprintf in the package). NULL silences them. */
void gsb_set_print(gsb_SimData* d, void (*print)(const char*));

/* ---- draws */
void gsb_set_draw_mode(gsb_SimData* d, int mode, uint64_t seed);
/* unif must return values in the open interval (0,1); rpois must return 0 for mu <= 0 */
void gsb_set_draw_source(gsb_SimData* d, double (*unif)(void), double (*rpois)(double));

/* ---- setup: the payload side of gsc_load_data_files (core.c:7453-7496) */
/* n genotypes, chars = n x 2M in the reference's layout; names may be NULL. New group. */
gsb_GroupNum gsb_load_genotypes(gsb_SimData* d, unsigned int n, const char* chars, const char* const* names);
/* gsc_load_genotypefile_matrix (core.c:7085-7428) for a simulation whose marker names are known (gsb_set_marker_names):
 * the reference's own detection of cell style, header row, orientation and corner cell; the parsed block goes to
 * the device in one piece.  New group, or 0 with a NOTE. */
gsb_GroupNum gsb_load_genotype_file(gsb_SimData* d, const char* filename);
gsb_MapID gsb_load_map(gsb_SimData* d, unsigned int n_chr, const double* expected_n_crossovers,
                       const unsigned int* chr_offset, const double* dists, const unsigned int* marker_index,
                       const unsigned char* has_dists);
gsb_EffectID gsb_load_effects(gsb_SimData* d, const unsigned int* cumn_alleles, const char* allele,
                              const double* eff, const double* centre);
void gsb_delete_recombination_map(gsb_SimData* d, gsb_MapID which);   /* core.c:4768-4830 */
void gsb_delete_eff_set(gsb_SimData* d, gsb_EffectID which);          /* core.c:4700-4752 */

/* ---- data access (core.c:4190-4420) */
unsigned int gsb_get_n_genotypes(const gsb_SimData* d);
unsigned int gsb_get_n_markers(const gsb_SimData* d);
const char* const* gsb_get_marker_names(const gsb_SimData* d);            /* NULL until gsb_set_marker_names */
/* ids of the recombination maps / effect sets that exist, in loading order (what gsc_SimData keeps in
 * genome.map_ids / eff_set_ids, core.h:863-901); out may be NULL; returns how many there are */
unsigned int gsb_get_map_ids(const gsb_SimData* d, unsigned int cap, gsb_MapID* out);
unsigned int gsb_get_eff_set_ids(const gsb_SimData* d, unsigned int cap, gsb_EffectID* out);
unsigned int gsb_get_index_of_name(const gsb_SimData* d, const char* name);   /* gsc_get_index_of_name (core.c:2513-2538): first match, else GSB_NA_GLOBALX */
unsigned int gsb_get_current_id(const gsb_SimData* d);
unsigned int gsb_get_group_size(const gsb_SimData* d, gsb_GroupNum group);
/* storage-order export of a group (0 = everything); any output may be NULL. genotypes is
 * cap x 2M chars. Returns the number of members written. */
unsigned int gsb_get_group_data(gsb_SimData* d, gsb_GroupNum group, unsigned int cap, char* genotypes,
                                unsigned int* ids, unsigned int* pedigree1, unsigned int* pedigree2,
                                unsigned int* groups, unsigned int* global_indexes);
const char* gsb_get_group_member_name(const gsb_SimData* d, gsb_GroupNum group, unsigned int i);
gsb_GroupNum gsb_make_group_from(gsb_SimData* d, unsigned int n, const unsigned int* global_indexes); /* core.c:2810-2841 */
void gsb_delete_group(gsb_SimData* d, gsb_GroupNum group);            /* core.c:4658-4692 */
gsb_GroupNum gsb_combine_groups(gsb_SimData* d, unsigned int n, const gsb_GroupNum* groups);   /* core.c:2705-2742 */

/* ---- group-split utilities (core.c:2980-3784): metadata only.  `results` receives the new group
 * numbers (the smallest free ones, in order of first appearance), the return value is how many groups were made.
 * Families come out of a hash table, one probe per genotype (the reference searches the families found so far). */
unsigned int gsb_split_into_families(gsb_SimData* d, gsb_GroupNum group, unsigned int maxentries_results, gsb_GroupNum* results);               /* core.c:3207-3230 */
unsigned int gsb_split_into_halfsib_families(gsb_SimData* d, gsb_GroupNum group, int parent, unsigned int maxentries_results, gsb_GroupNum* results); /* core.c:3125-3153 */
unsigned int gsb_split_into_individuals(gsb_SimData* d, gsb_GroupNum group, unsigned int maxentries_results, gsb_GroupNum* results);            /* core.c:3264-3271 */
gsb_GroupNum gsb_split_evenly_into_two(gsb_SimData* d, gsb_GroupNum group);                                                                      /* core.c:3298-3334 */
unsigned int gsb_split_evenly_into_n(gsb_SimData* d, gsb_GroupNum group, unsigned int n, gsb_GroupNum* results);                                 /* core.c:3444-3484 */
unsigned int gsb_split_into_buckets(gsb_SimData* d, gsb_GroupNum group, unsigned int n, const unsigned int* counts, gsb_GroupNum* results);      /* core.c:3521-3560 */
gsb_GroupNum gsb_split_randomly_into_two(gsb_SimData* d, gsb_GroupNum group);                                                                    /* core.c:3580-3602 */
unsigned int gsb_split_randomly_into_n(gsb_SimData* d, gsb_GroupNum group, unsigned int n, gsb_GroupNum* results);                               /* core.c:3645-3667 */
unsigned int gsb_split_by_probabilities(gsb_SimData* d, gsb_GroupNum group, unsigned int n, const double* probs, gsb_GroupNum* results);         /* core.c:3720-3763 */

/* ---- find.crossovers (core.c:7534-7818, experimental upstream): which parent every marker of an
 * offspring came from, judged marker by marker (window 1) or over a window; `origins` holds gsb_get_n_markers ints.
 * The from-file form reads "offspring parent1 parent2" names per line and writes the reference's table. 0 = done. */
int gsb_calculate_min_recombinations(gsb_SimData* d, gsb_MapID mapid, unsigned int parent1_index, unsigned int p1num,
                                     unsigned int parent2_index, unsigned int p2num, unsigned int offspring_index,
                                     int window_size, int certain, int* origins);                         /* core.c:7534, 7643 */
int gsb_calculate_recombinations_from_file(gsb_SimData* d, const char* input_file, const char* output_file,
                                           int window_len, int certain);                                   /* core.c:7760-7818 */

/* ---- crossing (core.c:8483-9212); same argument order as the reference */
gsb_GroupNum gsb_make_random_crosses(gsb_SimData* d, gsb_GroupNum from_group, unsigned int n_crosses,
                                     unsigned int cap, gsb_MapID which_map, gsb_GenOptions g);
gsb_GroupNum gsb_make_random_crosses_between(gsb_SimData* d, gsb_GroupNum group1, gsb_GroupNum group2,
                                             unsigned int n_crosses, unsigned int cap1, unsigned int cap2,
                                             gsb_MapID map1, gsb_MapID map2, gsb_GenOptions g);
gsb_GroupNum gsb_make_targeted_crosses(gsb_SimData* d, unsigned int n_combinations,
                                       const unsigned int* first_parents, const unsigned int* second_parents,
                                       gsb_MapID map1, gsb_MapID map2, gsb_GenOptions g);
gsb_GroupNum gsb_self_n_times(gsb_SimData* d, unsigned int n, gsb_GroupNum group, gsb_MapID which_map, gsb_GenOptions g);
gsb_GroupNum gsb_make_doubled_haploids(gsb_SimData* d, gsb_GroupNum group, gsb_MapID which_map, gsb_GenOptions g);
gsb_GroupNum gsb_make_clones(gsb_SimData* d, gsb_GroupNum group, _Bool inherit_names, gsb_GenOptions g);
gsb_GroupNum gsb_make_all_unidirectional_crosses(gsb_SimData* d, gsb_GroupNum from_group, gsb_MapID map, gsb_GenOptions g);
/* crossing plans from files of genotype names (core.c:9248-9313, 9347-9444) */
gsb_GroupNum gsb_make_crosses_from_file(gsb_SimData* d, const char* input_file, gsb_MapID map1, gsb_MapID map2, gsb_GenOptions g);
gsb_GroupNum gsb_make_double_crosses_from_file(gsb_SimData* d, const char* input_file, gsb_MapID map1, gsb_MapID map2, gsb_GenOptions g);

/* ---- evaluation (core.c:9463-10090) */
gsb_DecimalMatrix gsb_calculate_bvs(gsb_SimData* d, gsb_GroupNum group, gsb_EffectID effID);
unsigned int gsb_get_group_bvs(gsb_SimData* d, gsb_GroupNum group, gsb_EffectID effID, unsigned int group_size, double* output);
gsb_DecimalMatrix gsb_calculate_allele_counts(gsb_SimData* d, gsb_GroupNum group, char allele);
gsb_DecimalMatrix gsb_calculate_local_bvs(gsb_SimData* d, gsb_GroupNum group, gsb_MarkerBlocks b, gsb_EffectID effID);
gsb_MarkerBlocks gsb_create_evenlength_blocks_each_chr(const gsb_SimData* d, gsb_MapID mapid, unsigned int n);
/* gsc_load_blocks (core.c:9850-9918): one block per row after the header row, `chrom pos name class m1;m2;...`;
 * names the genome does not know are skipped.  Needs the marker names (the reference keeps them in its
 * KnownGenome; here the loader hands them over once): names[n_markers], copied. */
void gsb_set_marker_names(gsb_SimData* d, const char* const* names);
gsb_MarkerBlocks gsb_load_blocks(const gsb_SimData* d, const char* block_file);
/* optimal haplotypes / breeding values (core.c:10112-10392); opt_haplotype needs n_markers + 1 chars */
void gsb_calculate_optimal_haplotype(const gsb_SimData* d, gsb_EffectID effID, char symbol_na, char* opt_haplotype);
double gsb_calculate_optimal_bv(const gsb_SimData* d, gsb_EffectID effID);
double gsb_calculate_minimal_bv(const gsb_SimData* d, gsb_EffectID effID);
void gsb_calculate_optimal_possible_haplotype(gsb_SimData* d, gsb_GroupNum group, gsb_EffectID effID, char symbol_na, char* opt_haplotype);
double gsb_calculate_optimal_possible_bv(gsb_SimData* d, gsb_GroupNum group, gsb_EffectID effID);
gsb_GroupNum gsb_split_by_bv(gsb_SimData* d, gsb_GroupNum group, gsb_EffectID effID, unsigned int top_n, _Bool lowIsBest);
void gsb_delete_dmatrix(gsb_DecimalMatrix* m);          /* core.c:4628-4641 */
void gsb_delete_markerblocks(gsb_MarkerBlocks* b);      /* core.c:4757-4766 */

/* ---- savers (core.c:10449-10663): the reference's text formats byte for byte; genotypes leave the store
 *      1000 genotypes (or a block of markers) at a time.  The crossing functions above also write the
 *      save-as-you-go files of gsc_GenOptions (`{prefix}-pedigree.txt`, `-bv.txt`, `-genotype.txt`,
 *      core.c:8058-8163), and honour will_save_to_simdata = FALSE after writing them. */
void gsb_save_genotypes(const char* fname, gsb_SimData* d, gsb_GroupNum groupID, _Bool markers_as_rows);
void gsb_save_allele_counts(const char* fname, gsb_SimData* d, gsb_GroupNum groupID, char allele, _Bool markers_as_rows);
void gsb_save_pedigrees(const char* fname, gsb_SimData* d, gsb_GroupNum groupID, _Bool full_pedigree);
void gsb_save_bvs(const char* fname, gsb_SimData* d, gsb_GroupNum groupID, gsb_EffectID effID);
void gsb_save_local_bvs(const char* fname, gsb_SimData* d, gsb_GroupNum groupID, gsb_MarkerBlocks b, gsb_EffectID effID, _Bool headers);

/* ---- recurrent selection with the generation sharded over the GPUs of a box:
 *      one SimData per GPU, each holding a contiguous index range of every generation
 *      (gsb200_shard_range); GSB_DRAWS_NATIVE only.  Set-up: gsb_shard_init on every rank with the
 *      same sizes, gsb_shard_handle, hand all handles (rank order, GSB200_SHARD_HANDLE_BYTES each) to
 *      every rank, gsb_shard_connect.  All return 0 on success. */
int gsb_shard_init(gsb_SimData* d, unsigned int rank, unsigned int world, unsigned long long max_generation, unsigned int max_pool);
int gsb_shard_handle(gsb_SimData* d, void* handle);
int gsb_shard_connect(gsb_SimData* d, const void* handles);
/* gsc_split_by_bv (core.c:9463-9517) + gsc_make_random_crosses (core.c:8483-8532, cap 0, family_size 1) on
 * the whole sharded generation: `group` = this rank's slice of a generation of n_total (members in global
 * index order); the global top_n by GEBV are selected, exchanged, and this rank's slice of the
 * n_offspring_total random crosses among them becomes a new group (returned).  Ids, names and pedigrees
 * are those one process would have given offspring olo..ohi-1; every rank advances current_id by
 * n_offspring_total.  Nothing but row lists, ids and (for pedigrees) the picks crosses PCIe. */
gsb_GroupNum gsb_shard_select_and_cross(gsb_SimData* d, gsb_GroupNum group, unsigned long long n_total, gsb_EffectID effID,
                                        unsigned int top_n, _Bool lowIsBest, unsigned long long n_offspring_total,
                                        gsb_MapID which_map, gsb_GenOptions g);
/* The call above returns while the offspring are still being spliced (everything later on this SimData is
 * ordered behind them); this waits for them and reports a failed exchange or splice.  The next
 * gsb_shard_select_and_cross only POLLS for such a failure (it queues its own work behind the generation in
 * flight), gsb_delete_simdata waits.  0 on success. */
int gsb_shard_finish(gsb_SimData* d);
unsigned int gsb_shard_get_local_bvs(gsb_SimData* d, unsigned int n_local, double* output);
unsigned int gsb_shard_last_selection(gsb_SimData* d, unsigned long long n_total, double* bvs_out, unsigned int top_n, unsigned int* selected_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
