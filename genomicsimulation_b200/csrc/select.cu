// select.cu — K9: the ranking step of gsc_split_by_bv (core.c:9489-9508) as a radix SELECT.
//
// The reference sorts the whole group by breeding value (qsort of N pointers, core.c:9496) to take
// the first top_n.  Selection needs the SET, not the order: six histogram passes over the fp64
// keys (11 + 11 + 11 + 11 + 11 + 9 bits, most significant first) find the key T of the top_n-th
// best individual and how many of the individuals tied at T still fit; one more sweep writes the
// selected positions in ascending order — everything strictly better than T plus the first ties
// by position (the reference leaves tie order to qsort, core.c:1305-1306; here the rule is
// deterministic, so that every GPU of a sharded generation selects the same set).
//
// The first three passes are one kernel each over all the values: blocks histogram their slice in shared
// memory, add it to the global histogram, and the last block to finish scans the 2048 bins and extends the
// key prefix.  The third also writes out the keys that match the 22 bits known by then (a few hundred of a
// million for breeding values); the last three digits are then resolved over those candidates by ONE block
// (select_resolve_kernel) instead of three more sweeps — the selection is launch-latency-bound (3-9 us per
// launch), so six launches instead of nine is a third of its time.  The values stay where they are
// (8 bytes per individual, L2-resident at any realistic N).
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace gsb {

namespace {

constexpr int kPasses = 6;
constexpr int kBins = 2048;
constexpr int kHistThreads = 256;
constexpr int kItems = 8;                          // consecutive values per thread in the sweeps
constexpr int kSweepThreads = 256;
constexpr uint32_t kTile = kItems * kSweepThreads;

__host__ __device__ constexpr int pass_width(int pass) { return pass < 5 ? 11 : 9; }
__host__ __device__ constexpr int pass_shift(int pass) { return pass < 5 ? 64 - 11 * (pass + 1) : 0; }

// order-preserving map of fp64 to u64; the BEST value gets the SMALLEST key
__device__ __forceinline__ unsigned long long rank_key(double v, int low_is_best) {
    unsigned long long k = (unsigned long long) __double_as_longlong(v);
    if ((k << 1) == 0ull) k = 0ull;                // -0.0 and +0.0 compare equal in the reference's comparators (core.c:1308-1349): one key
    k = (k >> 63) ? ~k : (k | 0x8000000000000000ull);
    return low_is_best ? k : ~k;
}

__global__ void __launch_bounds__(kHistThreads) select_hist_kernel(const double* __restrict__ values, uint32_t n, int low_is_best,
                                                                   int pass, uint32_t top_n, SelectState* state,
                                                                   uint32_t* __restrict__ hist /* [kBins], zeroed */,
                                                                   uint32_t own_lo_init,
                                                                   unsigned long long* __restrict__ cand /* pass 2: the keys matching the prefix so far */,
                                                                   uint32_t* __restrict__ n_cand) {
    __shared__ uint32_t s_hist[kBins];
    __shared__ uint32_t s_scan[kHistThreads];
    __shared__ bool s_last;
    const int width = pass_width(pass), shift = pass_shift(pass);
    const uint32_t bins = 1u << width;
    for (uint32_t b = threadIdx.x; b < bins; b += kHistThreads) s_hist[b] = 0;
    __syncthreads();
    const unsigned long long prefix = pass ? state->prefix : 0ull;
    for (uint32_t i = blockIdx.x * kHistThreads + threadIdx.x; i < n; i += gridDim.x * kHistThreads) {
        const unsigned long long k = rank_key(values[i], low_is_best);
        if (pass == 0 || (k >> (shift + width)) == prefix) {
            atomicAdd(&s_hist[(uint32_t) (k >> shift) & (bins - 1)], 1u);
            if (cand) cand[atomicAdd(n_cand, 1u)] = k;
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += kHistThreads) if (s_hist[b]) atomicAdd(&hist[b], s_hist[b]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&state->done_blocks, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // the last block: find the bin holding the `remaining`-th best key among those matching the prefix
    __threadfence();
    const uint32_t remaining = pass ? state->remaining : top_n;
    const uint32_t per = bins / kHistThreads;                  // 8 or 2 consecutive bins per thread
    uint32_t mine = 0;
    for (uint32_t j = 0; j < per; ++j) {
        const uint32_t v = __ldcg(&hist[threadIdx.x * per + j]);
        s_hist[threadIdx.x * per + j] = v;
        mine += v;
    }
    s_scan[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 1; o < kHistThreads; o <<= 1) {               // inclusive scan over the threads' sums
        const uint32_t v = (int) threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0u;
        __syncthreads();
        s_scan[threadIdx.x] += v;
        __syncthreads();
    }
    const uint32_t incl = s_scan[threadIdx.x], excl = incl - mine;
    if (excl < remaining && remaining <= incl) {               // exactly one thread
        uint32_t before = excl;
        for (uint32_t j = 0; j < per; ++j) {
            const uint32_t v = s_hist[threadIdx.x * per + j];
            if (remaining <= before + v) {
                state->prefix = (prefix << width) | (threadIdx.x * per + j);
                state->remaining = remaining - before;
                break;
            }
            before += v;
        }
        state->done_blocks = 0;
        if (pass == 0) { state->own_lo = own_lo_init; state->own_hi = top_n; }
    }
}

// Passes first_pass .. kPasses - 1 over the candidate keys, by one block: histogram in shared memory, scan, extend the
// prefix, three times.  Any number of candidates is handled (a generation of identical breeding values makes every key
// a candidate; the block then simply loops longer).
__global__ void __launch_bounds__(1024) select_resolve_kernel(const unsigned long long* __restrict__ cand, const uint32_t* __restrict__ n_cand_p,
                                                              int first_pass, SelectState* state) {
    __shared__ uint32_t s_hist[kBins];
    __shared__ uint32_t s_scan[1024];
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_remaining;
    const uint32_t n_cand = *n_cand_p;
    if (threadIdx.x == 0) { s_prefix = state->prefix; s_remaining = state->remaining; }
    __syncthreads();
    for (int pass = first_pass; pass < kPasses; ++pass) {
        const int width = pass_width(pass), shift = pass_shift(pass);
        const uint32_t bins = 1u << width;
        for (uint32_t b = threadIdx.x; b < bins; b += 1024) s_hist[b] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const uint32_t remaining = s_remaining;
        for (uint32_t i = threadIdx.x; i < n_cand; i += 1024) {
            const unsigned long long k = cand[i];
            if ((k >> (shift + width)) == prefix) atomicAdd(&s_hist[(uint32_t) (k >> shift) & (bins - 1)], 1u);
        }
        __syncthreads();
        const uint32_t per = (bins + 1023) / 1024;                 // 2 or 1 consecutive bins per thread
        uint32_t mine = 0;
        for (uint32_t j = 0; j < per; ++j) if (threadIdx.x * per + j < bins) mine += s_hist[threadIdx.x * per + j];
        s_scan[threadIdx.x] = mine;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {                       // inclusive scan over the threads' sums
            const uint32_t v = (int) threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0u;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const uint32_t incl = s_scan[threadIdx.x], excl = incl - mine;
        if (excl < remaining && remaining <= incl) {               // exactly one thread
            uint32_t before = excl;
            for (uint32_t j = 0; j < per; ++j) {
                const uint32_t v = s_hist[threadIdx.x * per + j];
                if (remaining <= before + v) {
                    s_prefix = (prefix << width) | (threadIdx.x * per + j);
                    s_remaining = remaining - before;
                    break;
                }
                before += v;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { state->prefix = s_prefix; state->remaining = s_remaining; }
}

// per tile of kTile consecutive positions: how many keys are better than T, how many equal it
__global__ void __launch_bounds__(kSweepThreads) select_count_kernel(const double* __restrict__ values, uint32_t n, int low_is_best,
                                                                     const SelectState* __restrict__ state, uint2* __restrict__ tile_counts) {
    __shared__ uint32_t s_less[kSweepThreads / 32], s_eq[kSweepThreads / 32];
    const unsigned long long T = state->prefix;
    const uint32_t base = blockIdx.x * kTile + threadIdx.x * kItems;
    uint32_t less = 0, eq = 0;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        if (base + j < n) {
            const unsigned long long k = rank_key(values[base + j], low_is_best);
            less += k < T;
            eq += k == T;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { less += __shfl_xor_sync(0xffffffffu, less, o); eq += __shfl_xor_sync(0xffffffffu, eq, o); }
    if ((threadIdx.x & 31) == 0) { s_less[threadIdx.x >> 5] = less; s_eq[threadIdx.x >> 5] = eq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0, b = 0;
        for (int w = 0; w < kSweepThreads / 32; ++w) { a += s_less[w]; b += s_eq[w]; }
        tile_counts[blockIdx.x] = make_uint2(a, b);
    }
}

// exclusive scan of the tile counts, in place, one block
__global__ void __launch_bounds__(1024) select_scan_kernel(uint2* __restrict__ tile_counts, uint32_t n_tiles) {
    __shared__ uint2 s_warp[32];
    __shared__ uint2 s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += 1024) {
        const uint32_t t = t0 + threadIdx.x;
        const uint2 v = t < n_tiles ? tile_counts[t] : make_uint2(0, 0);
        uint2 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, incl.x, o), b = __shfl_up_sync(0xffffffffu, incl.y, o);
            if ((int) lane >= o) { incl.x += a; incl.y += b; }
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint2 before = s_carry;
        for (uint32_t w = 0; w < warp; ++w) { before.x += s_warp[w].x; before.y += s_warp[w].y; }
        if (t < n_tiles) tile_counts[t] = make_uint2(before.x + incl.x - v.x, before.y + incl.y - v.y);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = make_uint2(before.x + incl.x, before.y + incl.y);
        __syncthreads();
    }
}

// selected positions, ascending: key < T, or key == T and fewer than `remaining` ties before it
__global__ void __launch_bounds__(kSweepThreads) select_scatter_kernel(const double* __restrict__ values, uint32_t n, int low_is_best,
                                                                       SelectState* __restrict__ state, const uint2* __restrict__ tile_before,
                                                                       uint32_t* __restrict__ sel, uint32_t own_lo_pos, uint32_t own_hi_pos) {
    __shared__ uint2 s_warp[kSweepThreads / 32];
    const unsigned long long T = state->prefix;
    const uint32_t ties = state->remaining;
    const uint32_t base = blockIdx.x * kTile + threadIdx.x * kItems;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long key[kItems];
    uint32_t less = 0, eq = 0;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        key[j] = base + j < n ? rank_key(values[base + j], low_is_best) : ~0ull;
        if (base + j < n) { less += key[j] < T; eq += key[j] == T; }
    }
    uint2 incl = make_uint2(less, eq);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, incl.x, o), b = __shfl_up_sync(0xffffffffu, incl.y, o);
        if ((int) lane >= o) { incl.x += a; incl.y += b; }
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint2 before = tile_before[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) { before.x += s_warp[w].x; before.y += s_warp[w].y; }
    uint32_t lb = before.x + incl.x - less, eb = before.y + incl.y - eq;      // better / tied keys at positions before base
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        const uint32_t i = base + j;
        if (i >= n) break;
        const uint32_t pos = lb + min(eb, ties);                // selected individuals before position i
        if (i == own_lo_pos) state->own_lo = pos;
        if (i == own_hi_pos) state->own_hi = pos;
        if (key[j] < T) { sel[pos] = i; ++lb; }
        else if (key[j] == T) { if (eb < ties) sel[pos] = i; ++eb; }
    }
}


// ---- the same selection as ONE cooperative kernel ------------------------------------------------------------------------
// The seven launches above are 3-9 us each and the work between them is a few microseconds: at 100 k - 1 M values the
// selection is launch-latency-bound.  Here one grid of at most one block per SM walks the same phases with grid-wide
// barriers in between: three histogram sweeps over the block's contiguous slice (every block then scans the 2048 global
// bins for itself: no "last block" and no second barrier per pass), the last three digits from the candidate keys (every
// block for itself while they are few), per-block counts, and the ordered scatter.  Same result, bit for bit.
constexpr int kFusedThreads = 512;
constexpr uint32_t kFusedMinPerBlock = 1024;       // values per block below which more blocks only add barrier cost
constexpr uint32_t kRedundantCand = 16384;         // up to here every block resolves the candidates itself

// s_hist[0 .. bins) holds counts; the bin in which the running count reaches `remaining`: prefix is extended by it and
// remaining reduced by the count before it.  All threads of the block call this; all return the same values.
__device__ __forceinline__ void fused_find_bin(const uint32_t* s_hist, uint32_t* s_warp, uint32_t* s_out, uint32_t bins, int width,
                                               unsigned long long& prefix, uint32_t& remaining) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (bins + kFusedThreads - 1) / kFusedThreads;              // 4 or 1 consecutive bins per thread
    uint32_t mine = 0;
    for (uint32_t j = 0; j < per; ++j) if (threadIdx.x * per + j < bins) mine += s_hist[threadIdx.x * per + j];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int) lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    for (uint32_t w = 0; w < warp; ++w) incl += s_warp[w];
    const uint32_t excl = incl - mine;
    if (excl < remaining && remaining <= incl) {               // exactly one thread
        uint32_t before = excl;
        for (uint32_t j = 0; j < per; ++j) {
            const uint32_t v = s_hist[threadIdx.x * per + j];
            if (remaining <= before + v) { s_out[0] = threadIdx.x * per + j; s_out[1] = before; break; }
            before += v;
        }
    }
    __syncthreads();
    prefix = (prefix << width) | s_out[0];
    remaining -= s_out[1];
    __syncthreads();                                           // s_hist, s_warp and s_out are free again
}

__global__ void __launch_bounds__(kFusedThreads) select_fused_kernel(const double* __restrict__ values, uint32_t n, int low_is_best, uint32_t top_n,
                                                                     SelectState* state, uint32_t* hist /* [3][kBins], zeroed */,
                                                                     uint32_t* n_cand /* zeroed */, unsigned long long* cand,
                                                                     uint2* block_counts /* [gridDim.x] */, uint32_t* __restrict__ sel,
                                                                     uint32_t own_lo_pos, uint32_t own_hi_pos) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_hist[kBins];
    __shared__ uint32_t s_warp[kFusedThreads / 32];
    __shared__ uint2 s_warp2[kFusedThreads / 32];
    __shared__ uint32_t s_out[2];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = min(n, blockIdx.x * chunk), hi = min(n, lo + chunk);      // this block's slice, the same in every phase
    unsigned long long prefix = 0;
    uint32_t remaining = top_n;

    constexpr int kSweeps = 3;
    for (int pass = 0; pass < kSweeps; ++pass) {
        const int width = pass_width(pass), shift = pass_shift(pass);
        const uint32_t bins = 1u << width;
        for (uint32_t b = threadIdx.x; b < bins; b += kFusedThreads) s_hist[b] = 0;
        __syncthreads();
        for (uint32_t i = lo + threadIdx.x; i < hi; i += kFusedThreads) {
            const unsigned long long k = rank_key(values[i], low_is_best);
            if (pass == 0 || (k >> (shift + width)) == prefix) {
                atomicAdd(&s_hist[(uint32_t) (k >> shift) & (bins - 1)], 1u);
                if (pass == kSweeps - 1) cand[atomicAdd(n_cand, 1u)] = k;
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < bins; b += kFusedThreads) if (s_hist[b]) atomicAdd(&hist[pass * kBins + b], s_hist[b]);
        grid.sync();
        for (uint32_t b = threadIdx.x; b < bins; b += kFusedThreads) s_hist[b] = __ldcg(&hist[pass * kBins + b]);
        __syncthreads();
        fused_find_bin(s_hist, s_warp, s_out, bins, width, prefix, remaining);
    }

    // the last three digits, over the keys that match the 33 bits known by now
    const uint32_t nc = __ldcg(n_cand);
    const bool everyone = nc <= kRedundantCand;               // (the same answer in every block)
    if (everyone || blockIdx.x == 0) {
        for (int pass = kSweeps; pass < kPasses; ++pass) {
            const int width = pass_width(pass), shift = pass_shift(pass);
            const uint32_t bins = 1u << width;
            for (uint32_t b = threadIdx.x; b < bins; b += kFusedThreads) s_hist[b] = 0;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < nc; i += kFusedThreads) {
                const unsigned long long k = __ldcg(&cand[i]);
                if ((k >> (shift + width)) == prefix) atomicAdd(&s_hist[(uint32_t) (k >> shift) & (bins - 1)], 1u);
            }
            __syncthreads();
            fused_find_bin(s_hist, s_warp, s_out, bins, width, prefix, remaining);
        }
    }
    if (!everyone) {                                          // many ties (e.g. a generation of identical values): block 0 tells the others
        if (blockIdx.x == 0 && threadIdx.x == 0) { state->prefix = prefix; state->remaining = remaining; }
        grid.sync();
        prefix = __ldcg(&state->prefix);
        remaining = __ldcg(&state->remaining);
    }
    const unsigned long long T = prefix;
    const uint32_t ties = remaining;

    // how many better / tied keys my slice holds
    {
        uint32_t less = 0, eq = 0;
        for (uint32_t i = lo + threadIdx.x; i < hi; i += kFusedThreads) {
            const unsigned long long k = rank_key(values[i], low_is_best);
            less += k < T;
            eq += k == T;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { less += __shfl_xor_sync(0xffffffffu, less, o); eq += __shfl_xor_sync(0xffffffffu, eq, o); }
        if (lane == 0) s_warp2[warp] = make_uint2(less, eq);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint2 t = make_uint2(0, 0);
            for (int w = 0; w < kFusedThreads / 32; ++w) { t.x += s_warp2[w].x; t.y += s_warp2[w].y; }
            block_counts[blockIdx.x] = t;
            if (blockIdx.x == 0) {
                state->prefix = T; state->remaining = ties; state->done_blocks = 0;
                state->own_lo = own_lo_pos >= n ? top_n : 0u; state->own_hi = top_n;
            }
        }
    }
    grid.sync();
    uint32_t lb = 0, eb = 0;                                  // better / tied keys at positions before my slice
    if (warp == 0) {
        for (uint32_t j = lane; j < blockIdx.x; j += 32) { const uint2 c = __ldcg(&block_counts[j]); lb += c.x; eb += c.y; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lb += __shfl_xor_sync(0xffffffffu, lb, o); eb += __shfl_xor_sync(0xffffffffu, eb, o); }
        if (lane == 0) { s_out[0] = lb; s_out[1] = eb; }
    }
    __syncthreads();
    lb = s_out[0]; eb = s_out[1];
    __syncthreads();

    // selected positions, ascending: key < T, or key == T and fewer than `ties` ties before it
    for (uint32_t base = lo; base < hi; base += kFusedThreads) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < hi;
        const unsigned long long k = valid ? rank_key(values[i], low_is_best) : ~0ull;
        const bool is_less = valid && k < T, is_eq = valid && k == T;
        const uint32_t bl = __ballot_sync(0xffffffffu, is_less), be = __ballot_sync(0xffffffffu, is_eq);
        if (lane == 0) s_warp2[warp] = make_uint2(__popc(bl), __popc(be));
        __syncthreads();
        uint32_t l_before = lb + __popc(bl & ((1u << lane) - 1u)), e_before = eb + __popc(be & ((1u << lane) - 1u));
        uint32_t l_all = 0, e_all = 0;
#pragma unroll
        for (int w = 0; w < kFusedThreads / 32; ++w) {
            const uint2 c = s_warp2[w];
            if ((uint32_t) w < warp) { l_before += c.x; e_before += c.y; }
            l_all += c.x; e_all += c.y;
        }
        if (valid) {
            const uint32_t pos = l_before + min(e_before, ties);            // selected individuals before position i
            if (i == own_lo_pos) state->own_lo = pos;
            if (i == own_hi_pos) state->own_hi = pos;
            if (is_less) sel[pos] = i;
            else if (is_eq && e_before < ties) sel[pos] = i;
        }
        lb += l_all; eb += e_all;
        __syncthreads();
    }
}

}  // namespace

size_t select_scratch_bytes(uint32_t n) {
    const size_t tiles = ((size_t) n + kTile - 1) / kTile;
    return sizeof(SelectState) + (size_t) kPasses * kBins * sizeof(uint32_t) + tiles * sizeof(uint2) + 256 + 16 + (size_t) n * sizeof(unsigned long long);
}

// d_sel[0 .. top_n) <- positions (0 .. n-1) of the top_n best of d_values, ascending; `scratch` holds
// select_scratch_bytes(n) bytes and starts with the SelectState (left on the device: own_lo / own_hi are
// the ranks, within the selection, of positions own_lo_pos / own_hi_pos — the slice of the selection a
// shard owning [own_lo_pos, own_hi_pos) contributes).  Asynchronous on the context stream.
int select_top_n(gsb200_ctx* ctx, const double* d_values, uint32_t n, uint32_t top_n, int low_is_best, void* scratch,
                 uint32_t* d_sel, uint32_t own_lo_pos, uint32_t own_hi_pos) {
    GSB_REQUIRE(top_n >= 1 && top_n <= n, GSB200_ERR_ARG, "select: top_n %u out of range for %u values", top_n, n);
    SelectState* state = (SelectState*) scratch;
    uint32_t* hist = (uint32_t*) ((char*) scratch + sizeof(SelectState));
    uint2* tiles = (uint2*) (hist + (size_t) kPasses * kBins);
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    // behind the tile counts (rounded up to 256 bytes): the candidate counter, then the candidate keys
    char* after_tiles = (char*) scratch + ((sizeof(SelectState) + (size_t) kPasses * kBins * sizeof(uint32_t) + (size_t) n_tiles * sizeof(uint2) + 255) & ~(size_t) 255);
    uint32_t* n_cand = (uint32_t*) after_tiles;
    unsigned long long* cand = (unsigned long long*) (after_tiles + 16);
    // one cooperative kernel (default); GSB200_SELECT_PATH=multi keeps the seven-launch form for comparison
    static const bool multi = [] { const char* v = getenv("GSB200_SELECT_PATH"); return v && !strcmp(v, "multi"); }();
    if (!multi) {
        GSB_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(SelectState) + (size_t) 4 * kBins * sizeof(uint32_t), ctx->stream));
        const double* a_values = d_values;
        uint32_t a_n = n, a_top = top_n, a_lo = own_lo_pos, a_hi = own_hi_pos;
        int a_low = low_is_best;
        uint32_t* a_hist = hist;
        uint32_t* a_ncand = hist + 3 * kBins;                  // the sweeps use three of the six histograms: the counter and
        uint2* a_counts = (uint2*) (hist + 4 * kBins);         // the per-block counts live in the others
        void* args[] = {&a_values, &a_n, &a_low, &a_top, &state, &a_hist, &a_ncand, &cand, &a_counts, &d_sel, &a_lo, &a_hi};
        const uint32_t blocks = std::max(1u, std::min<uint32_t>((n + kFusedMinPerBlock - 1) / kFusedMinPerBlock, (uint32_t) ctx->sm_count));
        const cudaError_t le = cudaLaunchCooperativeKernel((const void*) select_fused_kernel, dim3(blocks), dim3(kFusedThreads), args, 0, ctx->stream);
        if (le == cudaSuccess) {
            ctx->launches += 1;
            return GSB200_OK;
        }
        // a device that cannot hold the grid at once right now (another client's kernels resident): the same selection as
        // separate launches below; anything else is an error
        if (le != cudaErrorCooperativeLaunchTooLarge && le != cudaErrorLaunchOutOfResources) GSB_CUDA_TRY(le);
        cudaGetLastError();
    }
    GSB_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(SelectState) + (size_t) kPasses * kBins * sizeof(uint32_t), ctx->stream));
    GSB_CUDA_TRY(cudaMemsetAsync(n_cand, 0, sizeof(uint32_t), ctx->stream));
    const uint32_t grid = (uint32_t) std::min<uint64_t>(((uint64_t) n + kHistThreads * 4 - 1) / (kHistThreads * 4), (uint64_t) ctx->sm_count * 8);
    constexpr int kSweeps = 3;                        // full sweeps; the rest of the digits come from the candidates
    for (int pass = 0; pass < kSweeps; ++pass)
        select_hist_kernel<<<std::max(grid, 1u), kHistThreads, 0, ctx->stream>>>(d_values, n, low_is_best, pass, top_n, state,
                                                                                hist + (size_t) pass * kBins, own_lo_pos >= n ? top_n : 0u,
                                                                                pass == kSweeps - 1 ? cand : nullptr, n_cand);
    select_resolve_kernel<<<1, 1024, 0, ctx->stream>>>(cand, n_cand, kSweeps, state);
    select_count_kernel<<<n_tiles, kSweepThreads, 0, ctx->stream>>>(d_values, n, low_is_best, state, tiles);
    select_scan_kernel<<<1, 1024, 0, ctx->stream>>>(tiles, n_tiles);
    select_scatter_kernel<<<n_tiles, kSweepThreads, 0, ctx->stream>>>(d_values, n, low_is_best, state, tiles, d_sel, own_lo_pos, own_hi_pos);
    ctx->launches += kSweeps + 1 + 3;
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB200_OK;
}

}  // namespace gsb
