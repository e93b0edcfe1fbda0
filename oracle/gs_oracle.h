/* oracle/gs_oracle.h — TEST INFRASTRUCTURE ONLY (the product never links this).
 *
 * A plain, sequential C restatement of the reference's meiosis + evaluation hot
 * path (genomicSimulation `src/sim-operations.c`), on flat arrays and in the
 * reference's own two-chars-per-marker genotype layout.  It exists so that the
 * CUDA path can be checked on machines where /root/reference (and therefore
 * oracle/_ref) is not available.  Parity status: PINNED — tests/test_oracle.py
 * checks it against the reference's own known answers (test-calculators.R:5-9,
 * 49-76, 104-141; test-group-utils.R:206-217; test-printers.R:67-111) via the
 * committed fixtures in tests/golden/, and — where oracle/_ref exists — against
 * the unmodified reference byte-for-byte on replayed draw tapes.
 *
 * Every function cites the reference file:line it follows ("core.c" =
 * src/sim-operations.c, "core.h" = src/sim-operations.h).
 */
#ifndef GSB200_GS_ORACLE_H
#define GSB200_GS_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gso_Sim gso_Sim;

/* GenOptions (core.h:629-669), flattened. */
typedef struct {
    unsigned family_size;
    int track_pedigree;
    int allocate_ids;
    int name_offspring;
    const char* name_prefix;
    int retain;                 /* will_save_to_simdata */
} gso_GenOptions;

gso_Sim* gso_new(unsigned n_markers);
void     gso_free(gso_Sim* s);

/* Append n genotypes (row-major n x 2M chars) as one new group; ids are handed
 * out sequentially like the loader does (core.c:7415-7421). Returns group. */
unsigned gso_add_genotypes(gso_Sim* s, unsigned n, const char* chars, const char* const* names);

/* Recombination map as concatenated linkage groups (core.h:678-765):
 * chr c covers positions [chr_offset[c], chr_offset[c+1]) of dists/marker_index.
 * has_dists[c]==0 reproduces the unlinked map's dists==NULL (core.c:6128-6152).
 * Returns map index. */
unsigned gso_add_map(gso_Sim* s, unsigned n_chr, const double* lambda, const unsigned* chr_offset,
                     const double* dists, const unsigned* marker_index, const unsigned char* has_dists);

/* Effect set in the reference's CSR layout (core.h:826-846); centre may be NULL. Returns index. */
unsigned gso_add_effects(gso_Sim* s, const unsigned* cumn_alleles, const char* allele,
                         const double* eff, const double* centre);

unsigned gso_n_genotypes(const gso_Sim* s);
unsigned gso_group_size(const gso_Sim* s, unsigned group);
unsigned gso_current_id(const gso_Sim* s);
/* storage-order export of a group (0 = all); any pointer may be NULL */
unsigned gso_export_group(const gso_Sim* s, unsigned group, unsigned cap, char* genotypes,
                          unsigned* ids, unsigned* ped1, unsigned* ped2, unsigned* groups,
                          unsigned* global_indexes);
const char* gso_member_name(const gso_Sim* s, unsigned group, unsigned i);

/* meiosis primitives (core.c:7850-7928, 7949-8033, 8047-8055) */
void gso_generate_gamete(const gso_Sim* s, const char* parent, char* out, unsigned map_index);
void gso_generate_doubled_haploid(const gso_Sim* s, const char* parent, char* out, unsigned map_index);
void gso_generate_clone(const gso_Sim* s, const char* parent, char* out);

/* crossing entry points (core.c:8483-9212); return new group number or 0 */
unsigned gso_make_random_crosses(gso_Sim* s, unsigned group, unsigned n_crosses, unsigned cap,
                                 unsigned map_index, gso_GenOptions g);
unsigned gso_make_random_crosses_between(gso_Sim* s, unsigned group1, unsigned group2, unsigned n_crosses,
                                         unsigned cap1, unsigned cap2, unsigned map1, unsigned map2,
                                         gso_GenOptions g);
unsigned gso_make_targeted_crosses(gso_Sim* s, unsigned n, const unsigned* first, const unsigned* second,
                                   unsigned map1, unsigned map2, gso_GenOptions g);
unsigned gso_make_all_unidirectional_crosses(gso_Sim* s, unsigned group, unsigned map_index, gso_GenOptions g);
unsigned gso_self_n_times(gso_Sim* s, unsigned n, unsigned group, unsigned map_index, gso_GenOptions g);
unsigned gso_make_doubled_haploids(gso_Sim* s, unsigned group, unsigned map_index, gso_GenOptions g);
unsigned gso_make_clones(gso_Sim* s, unsigned group, int inherit_names, gso_GenOptions g);

/* evaluation (core.c:9536-9667, 9697-9827, 9943-10090, 9463-9517) */
unsigned gso_calculate_bvs(const gso_Sim* s, unsigned group, unsigned eff_index, double* out);
unsigned gso_calculate_allele_counts(const gso_Sim* s, unsigned group, char allele, double* out);
/* blocks as CSR: offsets[n_blocks+1], markers[]; out is (2n) x n_blocks row-major; returns rows */
unsigned gso_calculate_local_bvs(const gso_Sim* s, unsigned group, unsigned n_blocks,
                                 const unsigned* offsets, const unsigned* markers,
                                 unsigned eff_index, double* out);
/* n blocks per chromosome; offsets needs n*n_chr+1 entries, markers needs n_markers entries.
 * Returns number of blocks. */
unsigned gso_create_evenlength_blocks_each_chr(const gso_Sim* s, unsigned map_index, unsigned n,
                                               unsigned* offsets, unsigned* markers);
unsigned gso_split_by_bv(gso_Sim* s, unsigned group, unsigned eff_index, unsigned top_n, int low_is_best);
unsigned gso_make_group_from(gso_Sim* s, unsigned n, const unsigned* global_indexes);
void     gso_delete_group(gso_Sim* s, unsigned group);

#ifdef __cplusplus
}
#endif
#endif
