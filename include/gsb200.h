/* gsb200.h — the C-ABI of the B200-native meiosis + genomic-evaluation engine.
 *
 * This is the drop-in boundary for genomicSimulation's hot path: plain C types,
 * opaque handle, `int` status (0 = ok) + gsb200_last_error().  Everything the
 * reference does inside `src/sim-operations.c` ("core.c") on the `gsc_AlleleMatrix`
 * payload for crossing and evaluation is reachable through these entry points; the
 * host-side mirror of the reference API (include/gsb200_gsc.h) and any future
 * patch of the reference's own core call nothing else.  No torch types, no C++.
 *
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *  - "row": index of one diploid genotype in the device-resident, bit-packed
 *    haplotype store (replaces `char* alleles[i]`, core.h:778-783).  Rows are
 *    chosen by the caller (the host core keeps the slot table).
 *  - pointer arguments are HOST pointers unless the function name ends in `_dev`.
 *  - all launches go to the context's stream; non-`_dev` functions synchronise
 *    before returning, `_dev` functions do not.
 *  - there is no CPU fallback: every function fails with GSB200_ERR_CUDA if the
 *    device is unusable.
 */
#ifndef GSB200_H
#define GSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define GSB200_ABI_VERSION 1

enum {
    GSB200_OK = 0,
    GSB200_ERR_CUDA = 1,        /* a CUDA runtime call failed / no device          */
    GSB200_ERR_ARG = 2,         /* invalid argument                                */
    GSB200_ERR_OOM = 3,         /* host or device allocation failed                */
    GSB200_ERR_UNSUPPORTED = 4, /* valid in the reference, not (yet) on this path  */
    GSB200_ERR_TAPE = 5         /* replay tape inconsistent with the request       */
};

typedef struct gsb200_ctx gsb200_ctx;

/* ------------------------------------------------------------------ context */

int  gsb200_abi_version(void);
const char* gsb200_last_error(void);             /* thread-local, never NULL */

/* One context per SimData (replaces the payload side of gsc_create_empty_simdata,
 * core.c:113-132, and gsc_delete_simdata, core.c:5047). */
int  gsb200_create(gsb200_ctx** out, int device_ordinal, uint32_t n_markers);
void gsb200_destroy(gsb200_ctx* ctx);
int  gsb200_synchronize(gsb200_ctx* ctx);
void* gsb200_stream(gsb200_ctx* ctx);            /* cudaStream_t, for timing with events */

/* ------------------------------------------------------------------ store
 * Bit-packed haplotype store. One row = 2 haplotypes x `planes` bit-planes x
 * `plane_words` 32-bit words; marker m is bit (m & 31) of word (m >> 5).  With
 * <= 2 distinct allele symbols at every marker there is one plane (1 bit per
 * haplotype per marker); a marker with more symbols (including '\0' = unknown,
 * core.c:75) widens the whole store to 2, 4 or 8 planes.  A per-marker symbol
 * dictionary maps codes back to the reference's chars. */
int  gsb200_store_reserve(gsb200_ctx* ctx, uint64_t n_rows);
uint64_t gsb200_store_capacity(const gsb200_ctx* ctx);
uint32_t gsb200_store_planes(const gsb200_ctx* ctx);
uint32_t gsb200_store_plane_words(const gsb200_ctx* ctx);
uint64_t gsb200_store_row_bytes(const gsb200_ctx* ctx);
void*    gsb200_store_device_ptr(gsb200_ctx* ctx);   /* base of row 0 (for NCCL / interop) */

/* chars (n x 2*n_markers, the reference's layout: alleles of marker m at [2m],[2m+1])
 * -> rows.  Replaces the founder fill gsc_helper_genotypecell_to_allelematrix
 * (core.c:6587-6674).  Extends the dictionary as needed. */
int  gsb200_store_upload_chars(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, const char* chars);
/* rows -> chars.  The save/see boundary: gsc_helper_output_genotypematrix_cell
 * (core.c:10951-10958), SXP_see_group_data "G" (r-wrappers.c:815-837). */
int  gsb200_store_download_chars(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, char* chars);
/* ... of markers first_marker .. first_marker + n_markers_out - 1 only (chars: n x 2*n_markers_out): the
 * markers-as-rows orientation of gsc_scaffold_save_genotype_info (core.c:10848-10907) walks one marker
 * across all genotypes, so the savers take the store a block of markers at a time. */
int  gsb200_store_download_chars_range(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t first_marker,
                                       uint32_t n_markers_out, char* chars);
/* Packed rows <-> a dense DEVICE buffer of n x gsb200_store_row_bytes() bytes (16-byte aligned):
 * the send/receive side of the per-generation parent-pool broadcast over NCCL (no reference
 * counterpart: the reference is single-process). `d_rows` is a DEVICE array; asynchronous on the
 * context stream, so that a collective enqueued on that stream needs no host synchronisation. */
int  gsb200_store_gather_rows_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, void* d_dense);
int  gsb200_store_scatter_rows_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, const void* d_dense);
/* gsc_change_allele_symbol (core.c:1029-1100) for one marker (marker == UINT32_MAX: all). */
int  gsb200_store_change_symbol(gsb200_ctx* ctx, uint32_t marker, char from, char to);
/* the whole dictionary: symbols[(m << planes) + code] and n_codes[m] (codes defined at marker m) */
int  gsb200_store_dictionary(const gsb200_ctx* ctx, char* symbols, uint16_t* n_codes);
/* dictionary of one marker: up to 2^planes symbols, returns how many are defined */
int  gsb200_store_symbols(const gsb200_ctx* ctx, uint32_t marker, char* out, uint32_t cap);

/* ------------------------------------------------------------------ maps
 * gsc_RecombinationMap (core.h:678-765) as concatenated linkage groups: chromosome c
 * covers positions [chr_offset[c], chr_offset[c+1]) of dists / marker_index;
 * lambda[c] = expected_n_crossovers; has_dists[c] == 0 reproduces dists == NULL of the
 * unlinked map (core.c:6128-6152).  `dists` must be the reference's own fp64 array
 * (core.c:5929): the splice rule `dists[i] > crossover` (core.c:7902-7903) is evaluated
 * on these exact doubles. */
int  gsb200_map_set(gsb200_ctx* ctx, uint32_t map_index, uint32_t n_chr, const double* lambda,
                    const uint32_t* chr_offset, const double* dists, const uint32_t* marker_index,
                    const uint8_t* has_dists);
int  gsb200_map_delete(gsb200_ctx* ctx, uint32_t map_index);
uint32_t gsb200_map_n_chr(const gsb200_ctx* ctx, uint32_t map_index);   /* 0 if absent */

/* ------------------------------------------------------------------ effects
 * gsc_MarkerEffects (core.h:826-846) in the reference's CSR layout; centre may be NULL. */
int  gsb200_effects_set(gsb200_ctx* ctx, uint32_t eff_index, const uint32_t* cumn_alleles,
                        const char* allele, const double* eff, const double* centre);
int  gsb200_effects_delete(gsb200_ctx* ctx, uint32_t eff_index);

/* ------------------------------------------------------------------ meiosis
 * Source of the random draws of gsc_generate_gamete (core.c:7873, 7889, 7896). */
enum { GSB200_DRAWS_PHILOX = 0, GSB200_DRAWS_TAPE = 1 };

typedef struct {
    uint32_t mode;
    /* PHILOX: counter-based stream keyed by (seed, stream) and indexed by
     * (first_unit + offspring, generation, gamete, chromosome): results do not depend on
     * how offspring are batched or sharded over GPUs. */
    uint64_t seed;
    uint64_t stream;
    uint64_t first_unit;
    /* TAPE (crossover replay): one entry per gamete-chromosome in the reference's own
     * consumption order — per offspring, per generation (selfing), per gamete, per
     * chromosome of that gamete's map: the Rf_rpois result, then that many unif_rand
     * positions (concatenated in `positions`, unsorted, as drawn), then the strand draw
     * already thresholded (unif_rand() > 0.5). */
    uint64_t n_entries;
    const uint32_t* n_crossovers;
    const uint8_t*  start_strand;
    uint64_t n_positions;
    const double*   positions;
} gsb200_draws;

/* gsc_helper_make_offspring_cross (core.c:8416-8423): offspring i = gamete(parent1[i], map1)
 * on strand 0, gamete(parent2[i], map2) on strand 1. */
int  gsb200_cross(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent1_rows, const uint32_t* parent2_rows,
                  uint32_t map1, uint32_t map2, const uint32_t* out_rows, const gsb200_draws* draws);
/* gsc_helper_make_offspring_self_n_times (core.c:8900-8926): n_generations rounds of two
 * gametes from the previous generation; only the last generation is stored. */
int  gsb200_self_n_times(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent_rows, uint32_t n_generations,
                         uint32_t map, const uint32_t* out_rows, const gsb200_draws* draws);
/* gsc_generate_doubled_haploid (core.c:7949-8033). */
int  gsb200_doubled_haploid(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent_rows, uint32_t map,
                            const uint32_t* out_rows, const gsb200_draws* draws);
/* gsc_generate_clone (core.c:8047-8055). */
int  gsb200_clone(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent_rows, const uint32_t* out_rows);

/* gsc_make_random_crosses / _between with no usage cap, parents chosen ON THE DEVICE
 * (gsc_helper_parentchooser_cross_randomly core.c:8364-8402, _between 8594-8631, pick rule
 * round(u * (n - 1)) redrawn on collision, core.c:8556-8581): n_crosses pairs, family_size
 * offspring each.  member_rows2 == NULL: both parents from group 1 and distinct.  out_rows ==
 * NULL: offspring go to rows first_out_row, first_out_row + 1, ...  picked1/2 (host, may be NULL)
 * receive, per offspring, the positions of its parents in member_rows1/2 — what the host needs for
 * pedigrees.  PHILOX draws only; nothing but the candidate lists and the picks crosses PCIe. */
int  gsb200_cross_random_pairs(gsb200_ctx* ctx, uint32_t n_crosses, uint32_t family_size,
                               const uint32_t* member_rows1, uint32_t n1, const uint32_t* member_rows2, uint32_t n2,
                               uint32_t map1, uint32_t map2, const uint32_t* out_rows, uint32_t first_out_row,
                               const gsb200_draws* draws, uint32_t* picked1, uint32_t* picked2);
/* The same in two halves, for callers with host work that needs only the picks: _begin returns as
 * soon as the picks are in picked1/2, with the offspring still being made; _end waits for them (and
 * redoes the batch on the general kernel in the rare case a gamete's plan overflowed).  No other
 * call may be made on the context between the two. */
int  gsb200_cross_random_pairs_begin(gsb200_ctx* ctx, uint32_t n_crosses, uint32_t family_size,
                               const uint32_t* member_rows1, uint32_t n1, const uint32_t* member_rows2, uint32_t n2,
                               uint32_t map1, uint32_t map2, const uint32_t* out_rows, uint32_t first_out_row,
                               const gsb200_draws* draws, uint32_t* picked1, uint32_t* picked2);
int  gsb200_cross_random_pairs_end(gsb200_ctx* ctx);

/* Device-pointer variants for callers that keep index arrays resident (benchmarks,
 * multi-GPU generation loops).  PHILOX draws only.  Asynchronous on the context stream. */
int  gsb200_cross_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_parent1_rows, const uint32_t* d_parent2_rows,
                      uint32_t map1, uint32_t map2, const uint32_t* d_out_rows, const gsb200_draws* draws);
/* gsb200_cross_dev plus the GEBVs of the offspring for effect set `eff_index` (d_bv[n], device), i.e.
 * make.random.crosses followed by see.GEBVs (wrap.c:1914 + 795-813) in one pass.  The GEBVs are NOT
 * computed by reading the offspring back: d_pool_rows[n_pool] (device) lists the distinct rows the
 * parents come from (the crossed group), a prefix table of effect sums is built per pool row, and
 * each gamete's value is assembled from the table at its crossover points while it is spliced.
 * Same value as gsb200_gebv_dev on the offspring up to fp64 rounding (both are exact fixed-point
 * sums converted once).  Falls back to the two separate passes when the store has several planes,
 * a map is not in marker order, or the tables would not fit.  A parent row outside the pool is
 * reported by gsb200_cross_dev_status. */
int  gsb200_cross_gebv_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_parent1_rows,
                           const uint32_t* d_parent2_rows, uint32_t map1, uint32_t map2,
                           const uint32_t* d_out_rows, const gsb200_draws* draws,
                           uint32_t n_pool, const uint32_t* d_pool_rows, uint32_t eff_index, double* d_bv);

/* Synchronises and reports whether any gsb200_cross_dev batch since the last check drew more
 * crossovers than its on-chip splice plan holds (probability < 1e-15 per gamete); such a batch
 * must be redone with gsb200_cross, which falls back to the general kernel by itself. */
int  gsb200_cross_dev_status(gsb200_ctx* ctx);

/* The native RNG's draws for `n_units` x `gametes_per_unit` gametes of map `map`, in tape
 * layout (for the KS tests against the reference's distributions).
 * n_crossovers/start_strand: n_units*gametes_per_unit*n_chr entries; positions: capacity
 * `positions_cap`, *n_positions receives the needed count. */
int  gsb200_export_draws(gsb200_ctx* ctx, uint32_t map, uint64_t n_units, uint32_t gametes_per_unit,
                         const gsb200_draws* philox, uint32_t* n_crossovers, uint8_t* start_strand,
                         double* positions, uint64_t positions_cap, uint64_t* n_positions);

/* ------------------------------------------------------------------ evaluation */

/* gsc_calculate_utility_bvs (core.c:9566-9618): out[i] = sum_m sum_{a in marker m}
 * count(a; row i, m) * eff(a, m)  -  sum_m centre[m].  fp64. */
int  gsb200_gebv(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t eff_index, double* out);
int  gsb200_gebv_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t eff_index, double* d_out);

/* gsc_calculate_utility_local_bvs (core.c:9979-10090). Blocks (gsc_MarkerBlocks,
 * core.h:515-528) as CSR: block b holds markers[offsets[b] .. offsets[b+1]).
 * out is row-major (2n) x n_blocks: rows 2i, 2i+1 = haplotypes of rows[i]. */
int  gsb200_local_gebv(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                       const uint32_t* block_offsets, const uint32_t* block_markers,
                       uint32_t eff_index, double* out);
/* The same for several effect sets in one pass over the genotypes (the reference computes one set
 * per call, core.c:9943-9961; see.local.GEBVs called once per set maps onto one call here):
 * outs[s] receives the matrix of eff_indexes[s]. */
int  gsb200_local_gebv_multi(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                             const uint32_t* block_offsets, const uint32_t* block_markers,
                             uint32_t n_effect_sets, const uint32_t* eff_indexes, double* const* outs);
/* ... with the output matrices left on the DEVICE (d_outs[s]: (2n) x n_blocks doubles each), for
 * consumers that keep working there; only for blocks that are contiguous, ascending,
 * non-overlapping marker ranges on a one-plane store (GSB200_ERR_UNSUPPORTED otherwise). */
int  gsb200_local_gebv_multi_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                                 const uint32_t* block_offsets, const uint32_t* block_markers,
                                 uint32_t n_effect_sets, const uint32_t* eff_indexes, double* const* d_outs);

/* gsc_calculate_allele_counts (core.c:9636-9667). out is row-major n x n_markers. */
enum { GSB200_COUNTS_F64 = 0, GSB200_COUNTS_I32 = 1, GSB200_COUNTS_U8 = 2 };
int  gsb200_allele_counts(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, char allele,
                          int out_kind, void* out);

/* Which alleles a set of genotypes carries: present[(m << planes) + code] = 1 if some strand of some
 * row has `code` at marker m.  The genotype scan of gsc_calculate_optimal_possible_haplotype /
 * gsc_calculate_optimal_possible_bv (core.c:10171-10222, 10285-10346) as an OR-reduction. */
int  gsb200_allele_presence(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint8_t* present);

/* The ranking step of gsc_split_by_bv (core.c:9489-9508) as a radix SELECT (no sort): positions
 * (0..n-1) of the `top_n` best values in ASCENDING POSITION order — the order the members of the
 * new group have in the reference too (storage order, core.c:9510-9514).  Everything strictly
 * better than the top_n-th best value is taken, then the ties at that value by lower position (the
 * reference leaves tie order to qsort, core.c:1305-1306). */
int  gsb200_top_n(gsb200_ctx* ctx, uint32_t n, const double* values, uint32_t top_n,
                  int low_is_best, uint32_t* out_positions);

/* ------------------------------------------------------------------ sharded generations
 * Recurrent selection with the generation sharded over the GPUs of one box (the
 * reference is single-process, so there is no counterpart beyond the calls named below): one
 * context — normally one process — per GPU holds a contiguous index range of the generation
 * (gsb200_shard_range).  All of a generation is enqueued on the context stream without any host
 * synchronisation; GEBVs and the selected parents travel by P2P stores over NVLink into buffers
 * the peers opened through CUDA IPC.
 *
 * Set-up: every rank creates its shard with the SAME sizes, exports a handle, the handles of all
 * ranks (in rank order) reach every rank by whatever the host has — torch.distributed, MPI, a
 * pipe — and gsb200_shard_connect opens them. */
typedef struct gsb200_shard gsb200_shard;
#define GSB200_SHARD_HANDLE_BYTES 128

/* [lo, hi) of rank's individuals in a generation of n_total: the first n_total % world ranks hold one more */
void gsb200_shard_range(uint64_t n_total, uint32_t rank, uint32_t world, uint64_t* lo, uint64_t* hi);
/* max_total: largest generation (all ranks together); max_pool: largest selection */
int  gsb200_shard_create(gsb200_ctx* ctx, uint32_t rank, uint32_t world, uint64_t max_total, uint32_t max_pool, gsb200_shard** out);
void gsb200_shard_destroy(gsb200_shard* shard);
int  gsb200_shard_handle(gsb200_shard* shard, void* handle /* [GSB200_SHARD_HANDLE_BYTES] */);
int  gsb200_shard_connect(gsb200_shard* shard, const void* handles /* world x GSB200_SHARD_HANDLE_BYTES, rank order */);

/* gsc_split_by_bv (core.c:9463-9517) over the whole sharded generation: GEBVs (effect set eff_index)
 * of this rank's n_local individuals (rows local_rows, global indexes lo .. hi-1 in that order) are
 * computed and exchanged, every rank selects the same global top_n (ties by lower global index), and
 * the selected individuals' packed rows land in every rank's parent pool in ascending global index.
 * local_ids (host, may be NULL): the individuals' PedigreeIDs, exchanged alongside so that every rank
 * can name the selected parents.  Every rank must make the same sequence of calls.  Asynchronous. */
int  gsb200_shard_select(gsb200_shard* shard, uint32_t n_local, const uint32_t* local_rows, const uint32_t* local_ids,
                         uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best);
/* ... local_rows == NULL: the individuals are rows first_local_row, first_local_row + 1, ... (no list crosses PCIe) */
int  gsb200_shard_select_run(gsb200_shard* shard, uint32_t n_local, const uint32_t* local_rows, uint32_t first_local_row,
                             const uint32_t* local_ids, uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best);
/* ... with the row list already on the device (it must stay valid until the call's work is done) */
int  gsb200_shard_select_dev(gsb200_shard* shard, uint32_t n_local, const uint32_t* d_local_rows, const uint32_t* local_ids,
                             uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best);
/* gsc_make_random_crosses (core.c:8483-8532, no usage cap, one offspring per cross) from the selected
 * pool: n offspring into out_rows of this rank's store; the pairs are drawn on the device with the pick
 * rule of core.c:8556-8581 from the stream position draws->first_unit (= the global index of this rank's
 * first offspring), so the generation does not depend on how it is sharded.  picked1/2 (host, may be
 * NULL): positions of the parents in the selection.  Synchronises. */
int  gsb200_shard_cross(gsb200_shard* shard, uint32_t n, uint32_t map, const uint32_t* out_rows, const gsb200_draws* draws,
                        uint32_t* picked1, uint32_t* picked2);
/* The same in three steps, for a host that has work of its own per generation (the host mirror's metadata):
 * _begin enqueues everything and returns at once; _picks waits only until the pairs have been drawn — long
 * before the offspring are spliced — and hands out the picks and the selected parents' ids (pinned buffers
 * owned by the shard, valid until the next _begin); _finish waits for the offspring and reports a splice-plan
 * overflow or a peer that never delivered.  The next gsb200_shard_select may be enqueued before _finish. */
int  gsb200_shard_cross_begin(gsb200_shard* shard, uint32_t n, uint32_t map, const uint32_t* out_rows, const gsb200_draws* draws, int want_picks);
/* ... out_rows == NULL: the offspring go to rows first_out_row, first_out_row + 1, ... */
int  gsb200_shard_cross_begin_run(gsb200_shard* shard, uint32_t n, uint32_t map, const uint32_t* out_rows, uint32_t first_out_row,
                                  const gsb200_draws* draws, int want_picks);
int  gsb200_shard_cross_picks(gsb200_shard* shard, const uint32_t** picked1, const uint32_t** picked2, const uint32_t** selected_ids);
int  gsb200_shard_finish(gsb200_shard* shard);
/* ... without waiting: reports a failure of an EARLIER generation once its flags have come down (they are sticky),
 * and asks for the current ones.  For a host that enqueues generation g + 1 while g is still being spliced. */
int  gsb200_shard_poll(gsb200_shard* shard);
/* ... asynchronous: d_out_rows NULL = rows first_out_row, first_out_row + 1, ...; d_picks (device, 2n) may be NULL */
int  gsb200_shard_cross_dev(gsb200_shard* shard, uint32_t n, uint32_t map, const uint32_t* d_out_rows, uint32_t first_out_row,
                            const gsb200_draws* draws, uint32_t* d_picks);
/* the last selection: global indexes (ascending), their ids, the GEBVs of the whole generation; any may be NULL. Synchronises. */
int  gsb200_shard_selected(gsb200_shard* shard, uint32_t* global_indexes, uint32_t* ids, double* all_bvs);
/* the GEBVs the last selection computed for THIS rank's n_local individuals, in their order.  After
 * gsb200_shard_select (host rows) they are already on their way down and this waits for that copy only. */
int  gsb200_shard_local_bvs(gsb200_shard* shard, uint32_t n_local, double* out);
uint64_t gsb200_shard_generation_size(const gsb200_shard* shard);   /* n_total of the last selection */
uint32_t gsb200_shard_pool_size(const gsb200_shard* shard);         /* top_n of the last selection */
void* gsb200_shard_pool_device_ptr(gsb200_shard* shard);   /* pool rows in store layout, selection order */
/* Phase timing with CUDA events on the context stream (off by default).  ms[6] of generation `epoch`
 * (1 = the first gsb200_shard_select since creation; the last 256 are kept): GEBV kernels | GEBV
 * exchange (push + wait for every rank) | selection | pool push | wait for every rank's pool rows |
 * pair draw + splice.  The two waits include waiting for the slowest rank.  Synchronises. */
int  gsb200_shard_set_timing(gsb200_shard* shard, int enabled);
uint32_t gsb200_shard_epoch(const gsb200_shard* shard);
int  gsb200_shard_phase_ms(gsb200_shard* shard, uint32_t epoch, float* ms /* [6] */);
/* synchronises; GSB200_ERR_CUDA if a peer failed to deliver its part of an exchange within ~20 s */
int  gsb200_shard_status(gsb200_shard* shard);

/* ------------------------------------------------------------------ instrumentation */
/* number of kernels launched by this context since creation */
uint64_t gsb200_kernel_launches(const gsb200_ctx* ctx);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
