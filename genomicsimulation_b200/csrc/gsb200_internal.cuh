// gsb200_internal.cuh — shared declarations of the CUDA side of libgsb200.so.
// Not part of the public boundary (that is include/gsb200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/gsb200.h"

namespace gsb {

void set_error(const char* fmt, ...);

#define GSB_CUDA_TRY(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            gsb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                           cudaGetErrorString(_e));                                         \
            return GSB200_ERR_CUDA;                                                         \
        }                                                                                   \
    } while (0)

#define GSB_TRY(expr)                                                                       \
    do {                                                                                    \
        int _s = (expr);                                                                    \
        if (_s != GSB200_OK) return _s;                                                     \
    } while (0)

#define GSB_REQUIRE(cond, code, ...)                                                        \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            gsb::set_error(__VA_ARGS__);                                                    \
            return (code);                                                                  \
        }                                                                                   \
    } while (0)

constexpr int kNumSMsB200 = 148;
constexpr uint32_t kPlaneAlignWords = 32;   // planes padded to 128 B

// A device buffer that only ever grows; contents are scratch.  With a stream (every scratch buffer of a
// context gets the context's) the memory comes from the stream-ordered allocator: growing a buffer then
// frees the old one IN STREAM ORDER instead of with cudaFree, which waits for the whole device — including,
// when several shards share a process, a peer's kernel that is spinning on OUR next push (csrc/shard.cu).
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    bool ordered = false;       // ptr came from cudaMallocAsync
    int ensure(size_t need);
    void release();
};

// One linkage group as the kernels see it (gsc_LinkageGroup, core.h:678-745).
struct ChrInfo {
    double   lambda;          // expected_n_crossovers
    double   exp_neg_lambda;  // exp(-lambda), the Poisson inversion's starting term
    uint32_t first;           // first marker index (contiguous groups)
    uint32_t len;             // markers in group
    uint32_t pos_off;         // offset of group in dists / marker_index
    uint32_t bucket_off;      // offset of group in the bucket index
    uint32_t bucket_shift;    // group has 2^shift buckets
    uint32_t searchable;      // 0: never switches strand (no dists, or all-NaN dists)
};

// Recombination map resident on the device (gsc_RecombinationMap, core.h:678-765).
// Chromosomes are in MAP order: the order the reference walks linkage groups and therefore
// the order of the draw tape.
struct DeviceMap {
    bool present = false;
    bool contiguous = false;        // every linkage group covers a contiguous marker range
    bool flat = false;              // ... the ranges are disjoint and together cover every marker:
                                    //     the warp-per-gamete splice applies
    bool covers_all = false;        // every marker belongs to some linkage group
    uint32_t n_chr = 0;
    uint32_t n_positions = 0;
    uint32_t n_chunks = 0;          // 32-position chunks over all chromosomes (general path)
    double lambda_sum = 0.0;
    double lambda_max = 0.0;
    ChrInfo*  d_chr = nullptr;      // [n_chr]
    double*   d_dists = nullptr;    // [n_positions] running maximum of the reference's dists (context.cu)
    uint32_t* d_marker_index = nullptr;
    uint32_t* d_bucket = nullptr;   // bucket index accelerating upper_bound(dists, x)
    // general path: chunk k covers positions [chunk_start[k], chunk_start[k]+32) of chromosome chunk_chr[k]
    uint32_t* d_chunk_chr = nullptr;
    uint32_t* d_chunk_start = nullptr;
    // markers no linkage group covers: gametes leave them at the reference's '\0' (core.c:75)
    std::vector<uint32_t> h_uncovered;
    uint32_t* d_uncovered = nullptr;
    void release();
};

// Effect set resident on the device (gsc_MarkerEffects, core.h:826-846).
struct DeviceEffects {
    bool present = false;
    std::vector<uint32_t> cumn;
    std::vector<uint8_t> allele;
    std::vector<double> eff;
    std::vector<double> centre;     // empty if none
    bool has_centre = false;
    double centre_sum = 0.0;        // sequential sum, as core.c:9604-9607
    // tables derived from the effect set AND the store's symbol dictionary:
    uint64_t built_for_dict_version = ~0ull;
    uint32_t built_for_planes = 0;
    double* d_value_all = nullptr;    // [M << planes]: sum of every entry whose allele == symbol(code)
    double* d_value_first = nullptr;  // [M << planes]: first such entry only (core.c:10019-10031)
    double* d_centre = nullptr;       // [M] or null
    double* d_delta = nullptr;        // planes == 1: value_all[m][1] - value_all[m][0]
    double  base_all = 0.0;           // planes == 1: sum_m 2 * value_all[m][0]
    void*   d_nonfinite = nullptr;    // NaN / Inf effect entries (marker, allele, effect), multiplied by their count like the reference does
    uint32_t n_nonfinite = 0;
    bool    all_finite = true;        // planes == 1: no NaN / Inf among the effects (else the fixed-point kernels do not apply)
    uint32_t* d_digit_frag = nullptr; // planes == 1: fixed-point digits of delta as mma B fragments
    uint32_t* d_digit_tc5 = nullptr;  // ... in the shared-memory layout of gebv_tc5_kernel
    double  digit_scale = 1.0;        // value of one fixed-point unit
    long long* d_bvq = nullptr;       // planes == 1: delta as one int64 per marker (fused cross + GEBV), padded to whole words
    double  bvq_scale = 1.0;          // value of one unit of d_bvq
    std::vector<double> h_value_first; // host copy of d_value_first (source of the local-GEBV digit fragments)
    uint64_t  local_key = 0;           // block definition d_local was built for
    uint32_t* d_local = nullptr;       // local GEBV: per-step digit fragments, then the per-block bases (fp64)
    uint32_t* d_local_tc5 = nullptr;   // ... the fragments in the shared-memory layout of local_tc5_kernel
    double    local_scale = 1.0;
    void release();
};

// State of one radix selection (select.cu), left on the device.
struct SelectState {
    unsigned long long prefix;     // key bits decided so far; after the last pass: the key T of the top_n-th best
    uint32_t remaining;            // still to take among the keys matching the prefix; after the last pass: ties at T taken
    uint32_t done_blocks;          // last-block-done counter of the pass in flight
    uint32_t own_lo, own_hi;       // slice [own_lo, own_hi) of the selection that lies in [own_lo_pos, own_hi_pos)
    uint32_t pad[2];
};

// Step schedule of the tensor-path local GEBV for the block definition last used.
struct LocalSchedule {
    bool valid = false, usable = false;
    uint64_t key = 0;
    std::vector<uint2> steps;           // (marker word, block)
    std::vector<uint32_t> block_first;  // first step of each non-empty block, then n_steps
    Scratch d_steps;
    Scratch d_plan;                     // local_tc5.cu: block ranges, runs, step records of the last launch
    Scratch d_bstream;                  // local_tc5.cu: the B digits of the pass's effect sets, interleaved per group of 8 steps
};

constexpr uint32_t kLocalTc5MaxSets = 10;   // effect sets per pass of the tcgen05 local-GEBV kernel (N = 8 * sets <= 80: three accumulators in TMEM)

}  // namespace gsb

struct gsb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t n_markers = 0;
    uint32_t planes = 1;
    uint32_t plane_words = 0;       // 32-bit words per plane (multiple of kPlaneAlignWords)
    uint64_t capacity = 0;          // rows
    uint32_t* d_store = nullptr;
    uint64_t launches = 0;
    // gsb200_cross_random_pairs_begin .. _end: what to redo should the splice plan have overflowed
    struct PendingCross {
        bool active = false;
        uint32_t n = 0, map1 = 0, map2 = 0;
        const uint32_t *d_p1 = nullptr, *d_p2 = nullptr, *d_out = nullptr;
        gsb200_draws draws;
    } pending_cross;
    int sm_count = gsb::kNumSMsB200;
    size_t l2_persist_max = 0;      // cudaDeviceProp::persistingL2CacheMaxSize
    size_t l2_window_max = 0;       // cudaDeviceProp::accessPolicyMaxWindowSize
    size_t l2_persist_set = 0;      // what cudaLimitPersistingL2CacheSize was last set to by this context
    int* d_flag = nullptr;          // kernel-side error flag (capacity overflow in a splice plan)

    // symbol dictionary: code -> char, per marker. host master copy + device mirror.
    std::vector<uint8_t> dict_sym;  // [M << planes]
    std::vector<uint16_t> dict_n;   // [M] number of codes defined (0..2^planes)
    uint8_t* d_dict_sym = nullptr;
    uint16_t* d_dict_n = nullptr;
    uint64_t dict_version = 0;
    bool dict_dirty = true;

    std::vector<gsb::DeviceMap> maps;
    std::vector<gsb::DeviceEffects> effects;

    gsb::Scratch s_rows[4];         // index arrays
    gsb::Scratch s_io;              // char staging / outputs
    gsb::Scratch s_plan[8];         // meiosis draw buffers
    gsb::Scratch s_eval[4];
    gsb::Scratch s_chain[2];        // ping-pong rows of chained selfing
    gsb::Scratch s_bv[3];           // fused cross + GEBV: parent prefix tables, row -> table slot, per-offspring sums
    gsb::LocalSchedule local_sched;
    gsb::Scratch s_select;          // radix selection: state, histograms, tile counts (select.cu)
    gsb::Scratch s_null;            // per marker: the code that decodes to '\0' (uncovered markers)
    uint64_t null_built_for = ~0ull;

    uint64_t row_words() const { return (uint64_t) 2 * planes * plane_words; }
};

namespace gsb {
int sync_dictionary(gsb200_ctx* ctx);                       // store.cu
int ensure_symbol(gsb200_ctx* ctx, uint32_t marker, uint8_t symbol, uint32_t* code);   // store.cu
int ensure_effect_tables(gsb200_ctx* ctx, uint32_t eff);    // evaluate.cu
void build_tc5_fragments(const std::vector<double>& delta, uint32_t n_markers, uint32_t plane_words, int exp2,
                         std::vector<uint32_t>& frag);     // gebv_tc5.cu
int gebv_tc5_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, const uint32_t* d_frag, unsigned long long* d_digits);
// local_tc5.cu
void build_local_tc5_fragments(const std::vector<double>& delta, const std::vector<uint2>& steps, const uint32_t* block_offsets,
                               const uint32_t* block_markers, int exp2, std::vector<uint32_t>& frag);
int local_tc5_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t n_blocks,
                     const std::vector<uint2>& steps, const std::vector<uint32_t>& block_first,
                     uint32_t n_sets, const uint32_t* const* d_b, const double* const* d_base, double* const* d_out, const double* scale);
int upload_u32(gsb200_ctx* ctx, Scratch& s, const uint32_t* host, size_t n, uint32_t** out);
int check_rows(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, const char* what);
// select.cu
size_t select_scratch_bytes(uint32_t n);
int select_top_n(gsb200_ctx* ctx, const double* d_values, uint32_t n, uint32_t top_n, int low_is_best, void* scratch,
                 uint32_t* d_sel, uint32_t own_lo_pos, uint32_t own_hi_pos);
// meiosis.cu: n random crosses among the n_pool rows of `pool_base` (packed rows in store layout, e.g. a
// shard's exchanged parent pool), pairs drawn on the device from the stream position draws->first_unit on;
// d_picks (device, 2n, may be null) receives the pool positions of both parents.  Asynchronous.
// A pool that is still arriving from the peers of a sharded generation (shard.cu): which rank owns which slots and where
// its "rows are here" flags live.  The splice then waits per parent, for the owner of that parent only.
struct PoolWait {
    const volatile uint32_t* flags;    // [world]: the epoch of rank r's rows present in this pool
    const uint32_t* bounds;            // device, [world + 1]: slots bounds[r] .. bounds[r + 1] - 1 are rank r's
    uint32_t epoch, world, rank;
    int* err;                          // raised when a peer does not show up in time
};
int pool_random_cross(gsb200_ctx* ctx, const uint32_t* pool_base, uint32_t n_pool, uint32_t n, uint32_t map,
                      const uint32_t* d_out_rows, uint32_t first_out_row, const gsb200_draws* draws, uint32_t* d_picks,
                      const char* what, uint32_t* h_picks = nullptr /* pinned, 2n: copy of d_picks */, cudaEvent_t picks_ready = nullptr,
                      const PoolWait* pw = nullptr, cudaStream_t side_stream = nullptr /* pairs and picks: beside what is queued in front */,
                      cudaEvent_t side_may_start = nullptr, cudaEvent_t pairs_done = nullptr);
}  // namespace gsb
