// evaluate.cu — genomic evaluation on the bit-packed store (K5 GEBV, K6 local GEBV,
// K7 allele counts; K9 ranking is select.cu).
//
// The reference compares allele CHARS against the effect table per marker
// (core.c:9587-9597, 10019-10031).  Here the per-marker symbol dictionary turns an effect set
// into tables indexed by the stored CODE:
//     value_all[m][code]   = sum of every effect entry of marker m whose allele is the symbol
//                            `code` decodes to            (whole-genome BVs count every match)
//     value_first[m][code] = the first such entry only    (local BVs stop at the first match)
// so that a genotype's contribution at marker m is value[m][code_hap0] + value[m][code_hap1].
// With one bit-plane (<= 2 symbols per marker) this collapses to
//     BV = sum_m 2*value[m][0]  +  sum_m (bit0_m + bit1_m) * (value[m][1] - value[m][0]).
// All arithmetic is fp64; only the summation ORDER differs from the reference's sequential
// sum (tolerance 1e-9 relative, BASELINE.json north_star).
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace gsb;

namespace {

// ------------------------------------------------------------------ GEBV, one plane
// BV_i = base + sum_m delta_m * (bit0_im + bit1_im) is a {0,1,2}-matrix times a vector.  Walking it
// marker by marker on the CUDA cores costs ~5 instructions (two of them FP64) per marker and
// individual — 4x more issue slots than the 1/4 byte per marker the HBM stream delivers.  So the
// product is done EXACTLY in integers on the tensor cores instead: delta_m is turned (on the
// host, once per effect set) into a 62-bit fixed-point number split into 8 balanced base-256
// digits, the digits are the 8 columns of the B operand of mma.m16n8k32 (u8 x s8 -> s32), the
// allele counts of 16 individuals x 32 markers are the A operand, and the eight int32 digit
// sums per individual are recombined in fp64 at the end.  Integer accumulation is exact and
// order-free, so marker segments can be summed with integer atomics and the result is still
// deterministic.
constexpr uint32_t kGebvMaxMarkers = 1u << 26;
constexpr uint32_t kChunkWords = 16;           // 512 markers = 16 mma = 64 bytes per strand row

__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(uint32_t* smem_dst, const uint32_t* gmem_src) {
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ... with an L2 eviction policy: lines this stream brings in are the first to go, so that it replaces ITSELF in L2 rather
// than the dirty lines the kernel in front (the splice) left there — evicting those costs a write-back per read miss
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_stream_hint(const uint4* p, unsigned long long pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}

// Persistent blocks, three per SM.  A block keeps the B fragments of ONE marker segment (<= 16
// chunks = 64 KB) in shared memory, loaded once, and its 8 warps pull groups of 16 individuals from
// the segment's counter: no per-chunk staging, no barrier after the first, no tail.
//   * lane (g, c) loads words 4c..4c+3 of a 16-word chunk of the two strands of individuals g and
//     g + 8 straight into registers (LDG.128; the four lanes of a row cover 64 contiguous bytes),
//     one chunk ahead — across group boundaries too — two chunks per loop trip so that no
//     register is ever copied;
//   * the two strands are ADDED before the tensor core sees them, so an A row is an individual,
//     not a strand, and one mma covers 16 individuals: with x = w0 ^ w1 and a = w0 & w1,
//     ye = (x & 0x55..) | ((a << 1) & 0xAA..) holds the allele count of every even-numbered marker
//     of the word as a 2-bit field, yo = ((x >> 1) & 0x55..) | (a & 0xAA..) that of every
//     odd-numbered one: 6 logic ops per word pair, then ONE op per A register,
//     `ye & (0x03030303 << 2p)`: byte j = count(marker 8j + 2p) at weight 4^p.  The weight is
//     divided out of the digits on the host (ensure_effect_tables);
//   * every lane uses the same masks, so they are immediates; mma (i, p) of a chunk contracts
//     k = 4c + j with marker bit 8j + 2p of word 4c + i and k = 16 + 4c + j with bit 8j + 2p + 1.
// bfrag[(chunk * 8 + i * 2 + p / 2) * 32 + lane] (uint4) = the B registers of `lane` for mma
// (i, p) and (i, p + 1), p even.
// What bounds it: the logic ops (alu pipe, one per 2 cycles and scheduler) and IMMA.16832 (~10
// cycles) do not overlap on sm_100a — their pipe-active percentages add up to ~95 %; without
// the row loads the kernel still takes 0.22 ms per 100 k x 50 k.
constexpr int kGebvWarps = 8;
constexpr int kGebvBlocksPerSm = 3;
constexpr uint32_t kGebvSegChunks = 16;        // B of a segment: 64 KB of shared memory; 8192 markers * 128 * 128 < 2^31
template <bool kEvictFirst>
__global__ void __launch_bounds__(kGebvWarps * 32, kGebvBlocksPerSm) gebv_counts_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ rows, uint32_t n, const uint4* __restrict__ bfrag,
        uint32_t n_groups, uint32_t n_segments, uint32_t seg_chunks, unsigned long long* __restrict__ digit_sums,
        unsigned int* __restrict__ next_group /* [n_segments], zeroed */) {
    extern __shared__ __align__(16) uint4 s_seg[];          // seg_chunks (rounded up to even) x 256
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t seg = blockIdx.x % n_segments;
    const uint32_t g = lane >> 2, c = lane & 3;
    const uint32_t ch_lo = seg * seg_chunks, ch_hi = min(ch_lo + seg_chunks, plane_words / kChunkWords);
    if (ch_lo >= ch_hi) return;                              // whole block
    const uint32_t segc = ch_hi - ch_lo, segc2 = (segc + 1) & ~1u;      // an odd segment gets a chunk of zero digits
    for (uint32_t i = threadIdx.x; i < segc2 * 256; i += kGebvWarps * 32) {
        if (i < segc * 256) cp_async16(reinterpret_cast<uint32_t*>(&s_seg[i]), reinterpret_cast<const uint32_t*>(bfrag + (size_t) ch_lo * 256 + i));
        else s_seg[i] = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // groups are handed out last-first: what a preceding kernel wrote last is still in L2
    auto take = [&]() {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(&next_group[seg], 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        return k < n_groups ? n_groups - 1 - k : 0xffffffffu;
    };
    const uint32_t strand1 = plane_words / 4;          // uint4 units
    struct Ptrs { const uint4* a; const uint4* b; };
    auto ptrs = [&](uint32_t group) {
        const uint32_t gr = min(group, n_groups - 1);
        Ptrs p;
        p.a = reinterpret_cast<const uint4*>(store + (uint64_t) rows[min(gr * 16 + g, n - 1)] * row_words) + ch_lo * 4 + c;
        p.b = reinterpret_cast<const uint4*>(store + (uint64_t) rows[min(gr * 16 + g + 8, n - 1)] * row_words) + ch_lo * 4 + c;
        return p;
    };
    uint32_t cur = take();
    if (cur == 0xffffffffu) return;
    uint32_t nxt = take();
    Ptrs lp = ptrs(cur), np = ptrs(nxt);
    struct Rows { uint4 a0, a1, b0, b1; };
    const unsigned long long pol = kEvictFirst ? l2_evict_first_policy() : 0ull;
    auto load = [&](const Ptrs& p, uint32_t ch, Rows& r) {
        const uint32_t o = min(ch, segc - 1) * 4;
        if (kEvictFirst) {
            r.a0 = ldg_stream_hint(p.a + o, pol); r.a1 = ldg_stream_hint(p.a + strand1 + o, pol);
            r.b0 = ldg_stream_hint(p.b + o, pol); r.b1 = ldg_stream_hint(p.b + strand1 + o, pol);
        } else {
            r.a0 = ldg_stream(p.a + o); r.a1 = ldg_stream(p.a + strand1 + o);
            r.b0 = ldg_stream(p.b + o); r.b1 = ldg_stream(p.b + strand1 + o);
        }
    };
    int acc0[4] = { 0, 0, 0, 0 }, acc1[4] = { 0, 0, 0, 0 };
    auto word = [&](uint32_t a_s0, uint32_t a_s1, uint32_t b_s0, uint32_t b_s1, const uint4* fw) {
        const uint32_t xa = a_s0 ^ a_s1, ca = a_s0 & a_s1, xb = b_s0 ^ b_s1, cb = b_s0 & b_s1;
        const uint32_t ae = (xa & 0x55555555u) | ((ca << 1) & 0xaaaaaaaau), ao = ((xa >> 1) & 0x55555555u) | (ca & 0xaaaaaaaau);
        const uint32_t be = (xb & 0x55555555u) | ((cb << 1) & 0xaaaaaaaau), bo = ((xb >> 1) & 0x55555555u) | (cb & 0xaaaaaaaau);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t m0 = 0x03030303u << (4 * q), m1 = 0x0c0c0c0cu << (4 * q);
            const uint4 fr = fw[q * 32];
            mma_u8s8(acc0, ae & m0, be & m0, ao & m0, bo & m0, fr.x, fr.y);
            mma_u8s8(acc1, ae & m1, be & m1, ao & m1, bo & m1, fr.z, fr.w);
        }
    };
    auto use = [&](uint32_t ch, const Rows& r) {
        const uint4* f = s_seg + ch * 256 + lane;
        word(r.a0.x, r.a1.x, r.b0.x, r.b1.x, f);
        word(r.a0.y, r.a1.y, r.b0.y, r.b1.y, f + 2 * 32);
        word(r.a0.z, r.a1.z, r.b0.z, r.b1.z, f + 4 * 32);
        word(r.a0.w, r.a1.w, r.b0.w, r.b1.w, f + 6 * 32);
    };
    Rows x, y;
    load(lp, 0, x);
    for (;;) {
        for (uint32_t ch = 0; ch < segc2; ch += 2) {             // two chunks per trip: no register shuffling
            load(lp, ch + 1, y);
            use(ch, x);
            if (ch + 2 < segc2) load(lp, ch + 2, x);
            else if (nxt != 0xffffffffu) load(np, 0, x);         // the next group's first chunk
            use(ch + 1, y);
        }
        // rows g / g + 8 of the accumulators = individuals ia / ib of `cur`; this lane: digits 2c, 2c + 1
        const uint32_t ia = cur * 16 + g, ib = ia + 8;
        if (ia < n) {
            atomicAdd(&digit_sums[(size_t) ia * 8 + 2 * c], (unsigned long long) (long long) (acc0[0] + acc1[0]));
            atomicAdd(&digit_sums[(size_t) ia * 8 + 2 * c + 1], (unsigned long long) (long long) (acc0[1] + acc1[1]));
        }
        if (ib < n) {
            atomicAdd(&digit_sums[(size_t) ib * 8 + 2 * c], (unsigned long long) (long long) (acc0[2] + acc1[2]));
            atomicAdd(&digit_sums[(size_t) ib * 8 + 2 * c + 1], (unsigned long long) (long long) (acc0[3] + acc1[3]));
        }
        if (nxt == 0xffffffffu) break;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc0[k] = acc1[k] = 0;
        cur = nxt;
        lp = np;
        nxt = take();
        np = ptrs(nxt);
    }
}

// BV_i = offset + scale * sum_j digit_sums[i][j] * 256^j, most significant digit first
__global__ void gebv_finish_kernel(const long long* __restrict__ digit_sums, uint32_t n, double scale, double offset, double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long* d = digit_sums + (size_t) i * 8;
    double v = (double) d[7];
#pragma unroll
    for (int j = 6; j >= 0; --j) v = v * 256.0 + (double) d[j];
    out[i] = offset + v * scale;
}

struct NonFiniteEntry { uint32_t marker; uint8_t allele; double eff; };

// ------------------------------------------------------------------ GEBV, any number of planes
// Warp per individual, lane per marker of a word; table lookups by code.
__global__ void gebv_generic_kernel(const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
                                    uint32_t planes, uint32_t n_markers, const uint32_t* __restrict__ rows, uint32_t n,
                                    const double* __restrict__ value, double offset, double* __restrict__ out,
                                    const uint8_t* __restrict__ dict_sym, const NonFiniteEntry* __restrict__ odd, uint32_t n_odd) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const uint32_t* row = store + (uint64_t) rows[warp] * row_words;
    double acc = 0.0;
    // NaN / Inf effect entries are kept out of the tables and multiplied by their count here, as the reference does
    // (count * eff, core.c:9590-9594): an allele the individual does NOT carry then contributes 0 * Inf = NaN
    for (uint32_t j = lane; j < n_odd; j += 32) {
        const uint32_t m = odd[j].marker;
        uint32_t c0 = 0, c1 = 0;
        for (uint32_t p = 0; p < planes; ++p) {
            c0 |= ((row[(uint64_t) p * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
            c1 |= ((row[(uint64_t) (planes + p) * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
        }
        const uint8_t* sym = dict_sym + ((size_t) m << planes);
        acc += (double) ((sym[c0] == odd[j].allele) + (sym[c1] == odd[j].allele)) * odd[j].eff;
    }
    for (uint32_t m = lane; m < n_markers; m += 32) {
        uint32_t c0 = 0, c1 = 0;
        for (uint32_t p = 0; p < planes; ++p) {
            c0 |= ((row[(uint64_t) p * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
            c1 |= ((row[(uint64_t) (planes + p) * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
        }
        const double* v = value + ((size_t) m << planes);
        acc += v[c0] + v[c1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[warp] = acc + offset;
}

// ------------------------------------------------------------------ local GEBV
// Warp per individual; for every block the lanes stride over the block's marker list.
__global__ void local_gebv_kernel(const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
                                  uint32_t planes, const uint32_t* __restrict__ rows, uint32_t n,
                                  uint32_t n_blocks, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ markers,
                                  const double* __restrict__ value_first, const double* __restrict__ centre,
                                  double* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const uint32_t* row = store + (uint64_t) rows[warp] * row_words;
    double* o1 = out + (size_t) (2 * (size_t) warp) * n_blocks;
    double* o2 = o1 + n_blocks;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        double a1 = 0.0, a2 = 0.0;
        for (uint32_t q = offsets[b] + lane; q < offsets[b + 1]; q += 32) {
            const uint32_t m = markers[q];
            uint32_t c0 = 0, c1 = 0;
            for (uint32_t p = 0; p < planes; ++p) {
                c0 |= ((row[(uint64_t) p * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
                c1 |= ((row[(uint64_t) (planes + p) * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
            }
            const double* v = value_first + ((size_t) m << planes);
            const double cm = centre ? centre[m] : 0.0;
            a1 += v[c0] - cm;
            a2 += v[c1] - cm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { o1[b] = a1; o2[b] = a2; }
    }
}

// ------------------------------------------------------------------ local GEBV, tensor path
// Blocks that are contiguous, ascending, non-overlapping marker ranges (what
// gsc_create_evenlength_blocks_each_chr produces on a map in marker order) on a one-plane store:
// the exact digit contraction of the GEBV kernel, per haplotype, with the accumulators flushed at
// every block boundary.  A rows are strands (the output is per haplotype, so the strands cannot
// be added first as gebv_counts_kernel does): row g = strand 0, row g + 8 = strand 1 of an
// individual; byte j of an A register = bit 8j + s of the word at weight 2^s (s = c for a0/a1,
// c + 4 for a2/a3).  A "step" is one 32-marker word of one block, with B digits zeroed outside
// the block (a word straddling two blocks is two steps).  A warp owns 16 individuals (two
// accumulator sets on the same B registers) x a range of whole blocks; its 32 strand rows are
// staged 64 bytes at a time with cp.async; the B fragments of 8-16 steps are shared by the
// block's four warps through shared memory.  kSets effect sets share each A register — with many
// effect sets this is the dense (strands x markers) x (markers x 8*sets) contraction of
// BASELINE.json configs[4].
constexpr uint32_t kLocalRowStride = 20;
constexpr uint32_t kLocalBufWords = 32 * kLocalRowStride;     // 16 individuals x 2 strands x one 64-byte chunk
constexpr int kLocalWarps = 4;
constexpr int kLocalMaxSets = 4;
constexpr uint32_t kLocalTileBytes = 16 * 1024;
constexpr uint32_t kLocalTc5MinSets = 5;         // from this many effect sets on the pass goes to local_tc5_kernel  // B fragments per shared-memory tile, all effect sets of the pass

struct LocalSetParams {
    const uint2* bfrag[kLocalMaxSets];     // [n_steps][32]
    const double* base[kLocalMaxSets];     // [n_blocks]
    double* out[kLocalMaxSets];            // [(2n) x n_blocks]
    double scale[kLocalMaxSets];
};

template <int kSets>
__global__ void __launch_bounds__(kLocalWarps * 32) local_gebv_digits_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ rows, uint32_t n, uint32_t n_blocks,
        const uint2* __restrict__ steps /* (word, block) */, const uint32_t* __restrict__ range_lo, uint32_t n_ranges,
        const LocalSetParams sp) {
    __shared__ __align__(16) uint32_t s_buf[kLocalWarps][2][kLocalBufWords];
    // B fragments of kLocalTile steps, shared by the block's warps (same range, 16 individuals each)
    constexpr uint32_t kLocalTile = kSets <= 2 ? 16 : 8;
    static_assert(2 * kSets * kLocalTile * 256 <= 2 * kLocalTileBytes, "tile too large");
    __shared__ __align__(16) uint4 s_bt[2][kSets][kLocalTile * 16];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_groups = (n + 15) / 16, group_blocks = (n_groups + kLocalWarps - 1) / kLocalWarps;
    const uint32_t range = blockIdx.x / group_blocks, group = (blockIdx.x % group_blocks) * kLocalWarps + warp;
    const uint32_t g = lane >> 2, c = lane & 3;
    const uint32_t s_lo = range_lo[range], s_hi = range_lo[range + 1];
    if (s_lo >= s_hi) return;                                // whole block
    // a warp past the last group walks the last rows (it must keep the barriers) and stores nothing

    // staging: 32 strand rows x 64 bytes per chunk; copy `it` moves 16 bytes of row it*8 + lane/4
    // (individual r & 15, strand r >> 4)
    const uint32_t my_row = rows[min(group * 16 + (lane & 15), n - 1)];
    uint32_t stage_row[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) stage_row[it] = __shfl_sync(0xffffffffu, my_row, (it * 8 + (lane >> 2)) & 15);
    const uint32_t* stage_base = store + (lane & 3) * 4;
    auto stage = [&](uint32_t chunk, uint32_t* buf) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const uint32_t r = it * 8 + (lane >> 2);
            cp_async16(buf + r * kLocalRowStride + (lane & 3) * 4,
                       stage_base + (uint64_t) stage_row[it] * row_words + (r >> 4) * plane_words + chunk * kChunkWords);
        }
        cp_async_commit();
    };

    const uint32_t k_lo = 0x01010101u << c, k_hi = k_lo << 4;
    const uint32_t n_chunks = plane_words / kChunkWords;
    // two accumulator sets on the same B registers: individuals ia = 16 * group + g and ib = ia + 8
    int acc[kSets][4], acc_b[kSets][4];
#pragma unroll
    for (int q = 0; q < kSets; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[q][k] = acc_b[q][k] = 0;

    const uint32_t ia = group * 16 + g, ib = ia + 8;
    const double w16 = exp2((double) (16 * c));
    // row g = strand 0 of the individual, row g + 8 = strand 1; this lane: digits 2c, 2c+1
    auto flush_one = [&](int (&a)[4], uint32_t ind, uint32_t block, int q) {
        double v0 = ((double) a[0] + 256.0 * (double) a[1]) * w16;
        double v1 = ((double) a[2] + 256.0 * (double) a[3]) * w16;
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        v0 += __shfl_xor_sync(0xffffffffu, v0, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        if (c == 0 && ind < n) {
            const double b = sp.base[q][block];
            sp.out[q][(size_t) (2 * (size_t) ind) * n_blocks + block] = b + v0 * sp.scale[q];
            sp.out[q][(size_t) (2 * (size_t) ind + 1) * n_blocks + block] = b + v1 * sp.scale[q];
        }
        a[0] = a[1] = a[2] = a[3] = 0;
    };
    auto flush = [&](uint32_t block) {
#pragma unroll
        for (int q = 0; q < kSets; ++q) { flush_one(acc[q], ia, block, q); flush_one(acc_b[q], ib, block, q); }
    };

    uint32_t cur_block = steps[s_lo].y, cur_chunk = 0xffffffffu, staged_next = 0xffffffffu, parity = 0;
    // one step, with its descriptor and B fragments already in registers
    auto step = [&](const uint2 st, const uint2 (&f)[kSets]) {
        if (st.y != cur_block) { flush(cur_block); cur_block = st.y; }
        const uint32_t chunk = st.x / kChunkWords;
        if (chunk != cur_chunk) {
            __syncwarp();                                   // everyone is done with the old buffers
            if (staged_next == chunk) { parity ^= 1; }
            else { cp_async_wait<0>(); parity = 0; stage(chunk, s_buf[warp][0]); }
            cur_chunk = chunk;
            if (chunk + 1 < n_chunks) { stage(chunk + 1, s_buf[warp][parity ^ 1]); staged_next = chunk + 1; cp_async_wait<1>(); }
            else { staged_next = 0xffffffffu; cp_async_wait<0>(); }
            __syncwarp();
        }
        const uint32_t* buf = s_buf[warp][parity] + g * kLocalRowStride + (st.x % kChunkWords);
        const uint32_t xa0 = buf[0], xa1 = buf[16 * kLocalRowStride];                      // individual ia: strand 0, 1
        const uint32_t xb0 = buf[8 * kLocalRowStride], xb1 = buf[24 * kLocalRowStride];    // individual ib
        const uint32_t a0 = xa0 & k_lo, a1 = xa1 & k_lo, a2 = xa0 & k_hi, a3 = xa1 & k_hi;
        const uint32_t b0 = xb0 & k_lo, b1 = xb1 & k_lo, b2 = xb0 & k_hi, b3 = xb1 & k_hi;
#pragma unroll
        for (int q = 0; q < kSets; ++q) {
            mma_u8s8(acc[q], a0, a1, a2, a3, f[q].x, f[q].y);
            mma_u8s8(acc_b[q], b0, b1, b2, b3, f[q].x, f[q].y);
        }
    };
    // B fragments go through shared memory a tile of kLocalTile steps at a time: read from global into
    // registers while the previous tile is being used, stored and published with one barrier per
    // tile.  (Read per step from L1/L2 straight before the mma, the fragment was half of all stall
    // samples.)  Step descriptors are requested a step ahead in two alternating register sets.
    constexpr uint32_t kPieces = kLocalTile * 16 / (kLocalWarps * 32);      // uint4 per thread, set and tile
    uint4 nxt[kSets][kPieces];
    auto tile_load = [&](uint32_t t0) {
#pragma unroll
        for (int q = 0; q < kSets; ++q)
#pragma unroll
            for (uint32_t k = 0; k < kPieces; ++k) {
                const uint32_t i = threadIdx.x + k * kLocalWarps * 32;                  // uint4 index in the tile
                const uint32_t sc = min(t0 + i / 16, s_hi - 1);
                nxt[q][k] = __ldg(reinterpret_cast<const uint4*>(sp.bfrag[q] + (size_t) sc * 32) + (i & 15));
            }
    };
    auto tile_store = [&](uint32_t buf) {
#pragma unroll
        for (int q = 0; q < kSets; ++q)
#pragma unroll
            for (uint32_t k = 0; k < kPieces; ++k) s_bt[buf][q][threadIdx.x + k * kLocalWarps * 32] = nxt[q][k];
    };
    tile_load(s_lo);
    tile_store(0);
    __syncthreads();
    uint2 st_a = __ldg(steps + s_lo), st_b;
    uint32_t buf = 0;
    for (uint32_t t0 = s_lo; t0 < s_hi; t0 += kLocalTile, buf ^= 1) {
        const uint32_t t1 = min(t0 + kLocalTile, s_hi);
        if (t1 < s_hi) tile_load(t1);
        for (uint32_t s = t0; s < t1; s += 2) {
            const uint2* fa = reinterpret_cast<const uint2*>(s_bt[buf][0]) + (s - t0) * 32 + lane;
            st_b = __ldg(steps + min(s + 1, s_hi - 1));
            uint2 f[kSets];
#pragma unroll
            for (int q = 0; q < kSets; ++q) f[q] = fa[q * kLocalTile * 32];
            step(st_a, f);
            if (s + 1 < t1) {
                st_a = __ldg(steps + min(s + 2, s_hi - 1));
#pragma unroll
                for (int q = 0; q < kSets; ++q) f[q] = fa[q * kLocalTile * 32 + 32];
                step(st_b, f);
            } else {
                st_a = st_b;                                     // odd tail (only at the end of the range)
            }
        }
        if (t1 < s_hi) tile_store(buf ^ 1);
        __syncthreads();
    }
    flush(cur_block);
    cp_async_wait<0>();
}

// ------------------------------------------------------------------ allele counts
template <typename T>
__global__ void allele_count_kernel(const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
                                    uint32_t planes, uint32_t n_markers, const uint32_t* __restrict__ rows, uint32_t n,
                                    const uint8_t* __restrict__ dict_sym, uint8_t allele, T* __restrict__ out) {
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) n * n_markers) return;
    const uint32_t i = (uint32_t) (tid / n_markers), m = (uint32_t) (tid % n_markers);
    const uint32_t* row = store + (uint64_t) rows[i] * row_words;
    uint32_t c0 = 0, c1 = 0;
    for (uint32_t p = 0; p < planes; ++p) {
        c0 |= ((row[(uint64_t) p * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
        c1 |= ((row[(uint64_t) (planes + p) * plane_words + (m >> 5)] >> (m & 31)) & 1u) << p;
    }
    const uint8_t* sym = dict_sym + ((size_t) m << planes);
    out[tid] = (T) ((sym[c0] == allele ? 1 : 0) + (sym[c1] == allele ? 1 : 0));
}

// ------------------------------------------------------------------ allele presence
// present[code][word] |= markers of `word` at which some strand of some row carries `code`: the
// OR-reduction behind gsc_calculate_optimal_possible_haplotype / _bv (core.c:10171-10222,
// 10285-10346), which scan every genotype's chars for the alleles a group contains.
__global__ void presence_kernel(const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words, uint32_t planes,
                                const uint32_t* __restrict__ rows, uint32_t n, uint32_t rows_per_slab, uint32_t* __restrict__ present) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= plane_words) return;
    const uint32_t lo = blockIdx.y * rows_per_slab, hi = min(lo + rows_per_slab, n);
    const uint32_t codes = 1u << planes;
    for (uint32_t c = 0; c < codes; ++c) {
        uint32_t acc = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t* row = store + (uint64_t) rows[i] * row_words + w;
#pragma unroll 1
            for (uint32_t h = 0; h < 2; ++h) {
                uint32_t m = 0xffffffffu;
                for (uint32_t p = 0; p < planes; ++p) {
                    const uint32_t x = row[(uint64_t) (h * planes + p) * plane_words];
                    m &= ((c >> p) & 1u) ? x : ~x;
                }
                acc |= m;
            }
        }
        if (acc) atomicOr(&present[(size_t) c * plane_words + w], acc);
    }
}

template <typename T>
int to_device_vec(const std::vector<T>& v, T** out) {
    if (*out) { cudaFree(*out); *out = nullptr; }
    GSB_CUDA_TRY(cudaMalloc((void**) out, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) GSB_CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GSB200_OK;
}

int get_effects(gsb200_ctx* ctx, uint32_t eff_index, DeviceEffects** out, const char* what) {
    GSB_REQUIRE(eff_index < ctx->effects.size() && ctx->effects[eff_index].present, GSB200_ERR_ARG,
                "%s: no effect set at index %u", what, eff_index);
    GSB_TRY(ensure_effect_tables(ctx, eff_index));
    *out = &ctx->effects[eff_index];
    return GSB200_OK;
}

}  // namespace

namespace gsb {

// (Re)build the code-indexed tables when the effect set is new or the dictionary changed.
int ensure_effect_tables(gsb200_ctx* ctx, uint32_t eff_index) {
    DeviceEffects& e = ctx->effects[eff_index];
    if (e.d_value_all && e.built_for_dict_version == ctx->dict_version && e.built_for_planes == ctx->planes) return GSB200_OK;
    const uint32_t M = ctx->n_markers, planes = ctx->planes, codes = 1u << planes;
    std::vector<double> all((size_t) M << planes, 0.0), first((size_t) M << planes, 0.0);
    for (uint32_t m = 0; m < M; ++m) {
        const uint32_t lo = m ? e.cumn[m - 1] : 0, hi = e.cumn[m];
        const uint8_t* sym = &ctx->dict_sym[(size_t) m << planes];
        // codes beyond dict_n[m] are never stored, except code 0 of a marker nobody uploaded:
        // it decodes to '\0', like the reference's zeroed slot (core.c:75)
        const uint32_t defined = std::max<uint32_t>(ctx->dict_n[m], 1);
        for (uint32_t c = 0; c < codes && c < defined; ++c) {
            double sum = 0.0;
            bool got = false;
            for (uint32_t x = lo; x < hi; ++x) {
                if (e.allele[x] != sym[c]) continue;
                if (std::isfinite(e.eff[x])) sum += e.eff[x];          // non-finite entries: see e.nonfinite below
                if (!got) { first[((size_t) m << planes) + c] = e.eff[x]; got = true; }
            }
            all[((size_t) m << planes) + c] = sum;
        }
    }
    // entries the whole-genome sum multiplies by their count (gebv_generic_kernel)
    std::vector<NonFiniteEntry> odd;
    for (uint32_t m = 0; m < M; ++m)
        for (uint32_t x = m ? e.cumn[m - 1] : 0; x < e.cumn[m]; ++x)
            if (!std::isfinite(e.eff[x])) odd.push_back(NonFiniteEntry{ m, e.allele[x], e.eff[x] });
    e.n_nonfinite = (uint32_t) odd.size();
    cudaStreamSynchronize(ctx->stream);
    {
        NonFiniteEntry* d_odd = (NonFiniteEntry*) e.d_nonfinite;
        GSB_TRY(to_device_vec(odd, &d_odd));
        e.d_nonfinite = d_odd;
    }
    GSB_TRY(to_device_vec(all, &e.d_value_all));
    GSB_TRY(to_device_vec(first, &e.d_value_first));
    e.h_value_first = first;
    if (e.has_centre) GSB_TRY(to_device_vec(e.centre, &e.d_centre));
    if (planes == 1) {
        std::vector<double> delta(M);
        double base = 0.0;
        for (uint32_t m = 0; m < M; ++m) {
            delta[m] = all[2 * (size_t) m + 1] - all[2 * (size_t) m];
            base += 2.0 * all[2 * (size_t) m];
        }
        GSB_TRY(to_device_vec(delta, &e.d_delta));
        e.base_all = base;
        // 62-bit fixed point, 8 balanced base-256 digits, in the mma B-fragment order of
        // gebv_counts_kernel: lane = 4 * digit + c holds, for mma (i, p) of a chunk, register `half`
        // whose byte j is the digit of marker bit 8j + 2p + half of word 4c + i.
        double maxabs = 0.0;
        e.all_finite = true;
        for (uint32_t m = 0; m < M; ++m) {
            maxabs = std::max(maxabs, std::fabs(delta[m]));
            e.all_finite = e.all_finite && std::isfinite(delta[m]) && std::isfinite(all[2 * (size_t) m]);
        }
        e.all_finite = e.all_finite && e.n_nonfinite == 0;
        int exp2 = 0;
        if (maxabs > 0.0 && std::isfinite(maxabs)) exp2 = 61 - std::ilogb(maxabs);
        // the kernel leaves the count of marker bit (8j + sft) of a word at weight 2^(sft & 6), so
        // the digits are those of round(delta * 2^(exp2 - (sft & 6))): 55..62 significant bits
        e.digit_scale = std::ldexp(1.0, -exp2);
        const uint32_t words = ctx->plane_words;
        std::vector<uint32_t> frag((size_t) words * 64, 0u);
        for (uint32_t m = 0; m < M; ++m) {
            const uint32_t w = m >> 5, bit = m & 31, j = bit >> 3, sft = bit & 7, p = sft >> 1, half = sft & 1;
            const uint32_t chunk = w / kChunkWords, wi = w % kChunkWords, cc = wi >> 2, i = wi & 3;
            long long q = std::isfinite(delta[m]) ? std::llrint(std::ldexp(delta[m], exp2 - (int) (sft & 6))) : 0;
            for (uint32_t dg = 0; dg < 8; ++dg) {
                const long long d8 = ((q + 128) & 255) - 128;
                q = (q - d8) >> 8;
                const uint32_t lane = 4 * dg + cc;
                frag[((((size_t) chunk * 8 + i * 2 + (p >> 1)) * 32 + lane) * 2 + (p & 1)) * 2 + half] |= ((uint32_t) (uint8_t) (int8_t) d8) << (8 * j);
            }
        }
        GSB_TRY(to_device_vec(frag, &e.d_digit_frag));
        build_tc5_fragments(delta, M, words, exp2, frag);
        GSB_TRY(to_device_vec(frag, &e.d_digit_tc5));
        // one int64 per marker for the fused cross + GEBV (meiosis.cu): sums over both strands of
        // every marker stay below 2^62
        int eq = 0;
        if (maxabs > 0.0 && std::isfinite(maxabs)) {
            int bits = 1;
            while ((1ull << bits) < 2ull * M) ++bits;
            eq = 60 - std::ilogb(maxabs) - bits;
        }
        e.bvq_scale = std::ldexp(1.0, -eq);
        std::vector<long long> bvq((size_t) words * 32, 0);
        for (uint32_t m = 0; m < M; ++m) bvq[m] = std::isfinite(delta[m]) ? std::llrint(std::ldexp(delta[m], eq)) : 0;
        GSB_TRY(to_device_vec(bvq, &e.d_bvq));
    }
    e.built_for_dict_version = ctx->dict_version;
    e.built_for_planes = planes;
    return GSB200_OK;
}

}  // namespace gsb

namespace {

int gebv_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t eff_index, double* d_out, const char* what) {
    DeviceEffects* e;
    GSB_TRY(get_effects(ctx, eff_index, &e, what));
    const double centre_sum = e->has_centre ? e->centre_sum : 0.0;
    // an effect set with NaN / Inf entries has no fixed-point image: it goes through the fp64 table kernel, which
    // propagates them into the breeding values as the reference's sum does (core.c:9587-9597)
    if (ctx->planes == 1 && ctx->n_markers <= kGebvMaxMarkers && e->all_finite) {
        const size_t counters = (size_t) (ctx->plane_words / kChunkWords + 1) * sizeof(unsigned int);
        GSB_TRY(ctx->s_eval[2].ensure((size_t) n * 8 * sizeof(long long) + counters));
        unsigned long long* d_digits = (unsigned long long*) ctx->s_eval[2].ptr;
        GSB_CUDA_TRY(cudaMemsetAsync(d_digits, 0, (size_t) n * 8 * sizeof(long long) + counters, ctx->stream));
        // the tcgen05 experiment (gebv_tc5.cu): same digits, same answer, slower
        static const bool use_tc5 = [] { const char* v = getenv("GSB200_GEBV_PATH"); return v && strcmp(v, "tc5") == 0; }();
        if (use_tc5) {
            GSB_TRY(gebv_tc5_launch(ctx, n, d_rows, e->d_digit_tc5, d_digits));
        } else {
            const uint32_t chunks = ctx->plane_words / kChunkWords, n_groups = (n + 15) / 16;
            const uint32_t want = (chunks + kGebvSegChunks - 1) / kGebvSegChunks, seg_chunks = (chunks + want - 1) / want;
            const uint32_t n_segments = (chunks + seg_chunks - 1) / seg_chunks;
            const uint64_t useful = (uint64_t) ((n_groups + kGebvWarps - 1) / kGebvWarps) * n_segments;
            const uint32_t grid = (uint32_t) std::min<uint64_t>(std::max<uint64_t>((uint64_t) ctx->sm_count * kGebvBlocksPerSm, n_segments), useful);
            // GSB200_GEBV_L2=evict_first: the row stream marked evict-first in L2 (A/B; see ldg_stream_hint)
            static const bool evict_first = [] { const char* v = getenv("GSB200_GEBV_L2"); return v && strcmp(v, "evict_first") == 0; }();
            auto* kernel = evict_first ? gebv_counts_kernel<true> : gebv_counts_kernel<false>;
            GSB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kGebvSegChunks * 4096));   // per device
            unsigned int* d_next = reinterpret_cast<unsigned int*>(d_digits + (size_t) n * 8);     // one counter per segment, zeroed above
            kernel<<<grid, kGebvWarps * 32, ((seg_chunks + 1) & ~1u) * 4096, ctx->stream>>>(
                ctx->d_store, ctx->row_words(), ctx->plane_words, d_rows, n, (const uint4*) e->d_digit_frag, n_groups, n_segments, seg_chunks,
                d_digits, d_next);
            ++ctx->launches;
            GSB_CUDA_TRY(cudaGetLastError());
        }
        gebv_finish_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const long long*) d_digits, n, e->digit_scale,
                                                                    e->base_all - centre_sum, d_out);
    } else {
        GSB_TRY(sync_dictionary(ctx));
        const uint64_t threads = (uint64_t) n * 32;
        gebv_generic_kernel<<<(unsigned) ((threads + 127) / 128), 128, 0, ctx->stream>>>(
            ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes, ctx->n_markers, d_rows, n,
            e->d_value_all, -centre_sum, d_out, ctx->d_dict_sym, (const NonFiniteEntry*) e->d_nonfinite, e->n_nonfinite);
    }
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB200_OK;
}

}  // namespace

extern "C" {

int gsb200_gebv_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t eff_index, double* d_out) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(d_rows && d_out, GSB200_ERR_ARG, "gsb200_gebv_dev: null array");
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    return gebv_launch(ctx, n, d_rows, eff_index, d_out, "gsb200_gebv_dev");
}

int gsb200_gebv(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t eff_index, double* out) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(out != nullptr, GSB200_ERR_ARG, "gsb200_gebv: null output");
    GSB_TRY(check_rows(ctx, n, rows, "gsb200_gebv"));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t* d_rows;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], rows, n, &d_rows));
    GSB_TRY(ctx->s_eval[0].ensure((size_t) n * sizeof(double)));
    GSB_TRY(gebv_launch(ctx, n, d_rows, eff_index, (double*) ctx->s_eval[0].ptr, "gsb200_gebv"));
    GSB_CUDA_TRY(cudaMemcpyAsync(out, ctx->s_eval[0].ptr, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

// Build the step schedule of the tensor path; false if the blocks do not qualify.
static bool local_schedule(const gsb200_ctx* ctx, uint32_t n_blocks, const uint32_t* off, const uint32_t* mk,
                           std::vector<uint2>& steps, std::vector<uint32_t>& step_block_first) {
    if (ctx->planes != 1) return false;
    uint32_t prev_end = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const uint32_t lo = off[b], hi = off[b + 1];
        if (lo == hi) continue;
        const uint32_t m0 = mk[lo], len = hi - lo;
        if (len > 32768 || m0 < prev_end) return false;
        for (uint32_t q = 1; q < len; ++q) if (mk[lo + q] != m0 + q) return false;
        prev_end = m0 + len;
        step_block_first.push_back((uint32_t) steps.size());
        for (uint32_t w = m0 >> 5; w <= (m0 + len - 1) >> 5; ++w) steps.push_back(make_uint2(w, b));
    }
    step_block_first.push_back((uint32_t) steps.size());
    return !steps.empty();
}

static int local_gebv_generic(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t n_blocks, const uint32_t* d_off,
                              const uint32_t* d_mk, DeviceEffects* e, double* out) {
    // output in slabs of individuals so that the device buffer stays bounded
    const size_t row_out = 2 * (size_t) n_blocks * sizeof(double);
    const uint32_t slab = (uint32_t) std::max<size_t>(1, std::min<size_t>(n, ((size_t) 1 << 30) / row_out));
    GSB_TRY(ctx->s_eval[1].ensure((size_t) slab * row_out));
    for (uint32_t done = 0; done < n; done += slab) {
        const uint32_t cnt = std::min(slab, n - done);
        const uint64_t threads = (uint64_t) cnt * 32;
        local_gebv_kernel<<<(unsigned) ((threads + 127) / 128), 128, 0, ctx->stream>>>(
            ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes, d_rows + done, cnt, n_blocks, d_off, d_mk,
            e->d_value_first, e->has_centre ? e->d_centre : nullptr, (double*) ctx->s_eval[1].ptr);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
        GSB_CUDA_TRY(cudaMemcpyAsync(out + (size_t) done * 2 * n_blocks, ctx->s_eval[1].ptr, (size_t) cnt * row_out,
                                     cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return GSB200_OK;
}

static int local_gebv_multi_impl(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                                 const uint32_t* block_offsets, const uint32_t* block_markers,
                                 uint32_t n_effect_sets, const uint32_t* eff_indexes, double* const* outs, bool device_out) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = device_out ? "gsb200_local_gebv_multi_dev" : "gsb200_local_gebv";
    if (n == 0 || n_blocks == 0 || n_effect_sets == 0) return GSB200_OK;
    GSB_REQUIRE(block_offsets && outs && eff_indexes, GSB200_ERR_ARG, "%s: null array", what);
    GSB_TRY(check_rows(ctx, n, rows, what));
    GSB_REQUIRE(block_offsets[0] == 0, GSB200_ERR_ARG, "%s: block_offsets[0] must be 0", what);
    for (uint32_t b = 0; b < n_blocks; ++b)
        GSB_REQUIRE(block_offsets[b] <= block_offsets[b + 1], GSB200_ERR_ARG, "%s: block_offsets not ascending", what);
    const uint32_t total = block_offsets[n_blocks];
    GSB_REQUIRE(total == 0 || block_markers, GSB200_ERR_ARG, "%s: null marker list", what);
    for (uint32_t q = 0; q < total; ++q)
        GSB_REQUIRE(block_markers[q] < ctx->n_markers, GSB200_ERR_ARG, "%s: marker index %u out of range", what, block_markers[q]);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<DeviceEffects*> effs(n_effect_sets);
    for (uint32_t q = 0; q < n_effect_sets; ++q) {
        GSB_REQUIRE(outs[q] != nullptr, GSB200_ERR_ARG, "%s: null output", what);
        GSB_TRY(get_effects(ctx, eff_indexes[q], &effs[q], what));
    }
    uint32_t* d_rows;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], rows, n, &d_rows));

    // The step schedule and each effect set's digit fragments depend only on the blocks (and the
    // dictionary): they are cached on the device under a hash of the block definition, so that a
    // breeding scheme evaluating the same blocks every generation pays for them once.
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](const uint32_t* p, size_t cnt) { for (size_t i = 0; i < cnt; ++i) { key ^= p[i]; key *= 1099511628211ull; } };
    mix(&n_blocks, 1);
    mix(block_offsets, (size_t) n_blocks + 1);
    mix(block_markers, total);
    key ^= ctx->dict_version * 0x9E3779B97F4A7C15ull;
    key ^= (uint64_t) ctx->planes << 56;
    LocalSchedule& sc = ctx->local_sched;
    if (sc.key != key || !sc.valid) {
        sc.steps.clear();
        sc.block_first.clear();
        sc.usable = local_schedule(ctx, n_blocks, block_offsets, block_markers, sc.steps, sc.block_first);
        if (sc.usable) {
            GSB_TRY(sc.d_steps.ensure(sc.steps.size() * sizeof(uint2)));
            GSB_CUDA_TRY(cudaMemcpyAsync(sc.d_steps.ptr, sc.steps.data(), sc.steps.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
            GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        sc.key = key;
        sc.valid = true;
    }
    bool finite = true;
    for (DeviceEffects* e : effs) finite = finite && e->all_finite;
    if (!sc.usable || !finite) {
        GSB_REQUIRE(!device_out, GSB200_ERR_UNSUPPORTED,
                    "%s: device output needs blocks that are contiguous, ascending, non-overlapping marker ranges on a one-plane store", what);
        uint32_t *d_off, *d_mk;
        GSB_TRY(upload_u32(ctx, ctx->s_rows[1], block_offsets, (size_t) n_blocks + 1, &d_off));
        GSB_TRY(upload_u32(ctx, ctx->s_rows[2], block_markers, total, &d_mk));
        for (uint32_t q = 0; q < n_effect_sets; ++q)
            GSB_TRY(local_gebv_generic(ctx, n, d_rows, n_blocks, d_off, d_mk, effs[q], outs[q]));
        return GSB200_OK;
    }

    // ---- tensor path
    const std::vector<uint2>& steps = sc.steps;
    const std::vector<uint32_t>& block_first = sc.block_first;
    const uint32_t M = ctx->n_markers, n_steps = (uint32_t) steps.size();
    // block ranges: tasks = (16 individuals) x (range of whole blocks), a few waves over the machine
    const uint32_t n_groups = (n + 15) / 16;
    uint32_t want_ranges = (uint32_t) std::max<uint64_t>(1, ((uint64_t) ctx->sm_count * 48 * 4 + n_groups - 1) / n_groups);
    want_ranges = std::min<uint32_t>(want_ranges, (uint32_t) block_first.size() - 1);
    std::vector<uint32_t> range_lo(1, 0);
    for (uint32_t r = 1; r < want_ranges; ++r) {
        const uint32_t target = (uint32_t) ((uint64_t) n_steps * r / want_ranges);
        auto it = std::lower_bound(block_first.begin(), block_first.end(), target);
        if (it != block_first.end() && *it > range_lo.back() && *it < n_steps) range_lo.push_back(*it);
    }
    range_lo.push_back(n_steps);
    const uint32_t n_ranges = (uint32_t) range_lo.size() - 1;
    uint32_t* d_ranges;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[1], range_lo.data(), range_lo.size(), &d_ranges));
    const uint2* d_steps = (const uint2*) sc.d_steps.ptr;
    const size_t frag_bytes = (size_t) n_steps * 32 * sizeof(uint2);

    // fragments + block bases of every effect set asked for (cached per set)
    for (uint32_t q = 0; q < n_effect_sets; ++q) {
        DeviceEffects* e = effs[q];
        if (e->local_key == key && e->d_local) continue;
        const std::vector<double>& vf = e->h_value_first;      // [M << 1]
        double maxabs = 0.0;
        for (uint32_t m = 0; m < M; ++m) maxabs = std::max(maxabs, std::fabs(vf[2 * (size_t) m + 1] - vf[2 * (size_t) m]));
        int exp2v = 0;
        if (maxabs > 0.0 && std::isfinite(maxabs)) exp2v = 61 - std::ilogb(maxabs);
        e->local_scale = std::ldexp(1.0, -exp2v);
        std::vector<uint32_t> frag((size_t) n_steps * 64 + 2 * (size_t) n_blocks, 0u);
        double* base = reinterpret_cast<double*>(frag.data() + (size_t) n_steps * 64);
        for (uint32_t st = 0; st < n_steps; ++st) {
            const uint32_t w = steps[st].x, b = steps[st].y;
            const uint32_t m_lo = block_markers[block_offsets[b]], m_hi = m_lo + (block_offsets[b + 1] - block_offsets[b]);
            for (uint32_t bit = 0; bit < 32; ++bit) {
                const uint32_t m = w * 32 + bit;
                if (m < m_lo || m >= m_hi) continue;
                const double d = vf[2 * (size_t) m + 1] - vf[2 * (size_t) m];
                const uint32_t j = bit >> 3, sft = bit & 7, cc = sft & 3, half = sft >> 2;
                long long qv = std::isfinite(d) ? std::llrint(std::ldexp(d, exp2v - (int) sft)) : 0;
                for (uint32_t dg = 0; dg < 8; ++dg) {
                    const long long d8 = ((qv + 128) & 255) - 128;
                    qv = (qv - d8) >> 8;
                    frag[((size_t) st * 32 + 4 * dg + cc) * 2 + half] |= ((uint32_t) (uint8_t) (int8_t) d8) << (8 * j);
                }
            }
        }
        for (uint32_t b = 0; b < n_blocks; ++b) {
            double acc = 0.0;
            for (uint32_t x = block_offsets[b]; x < block_offsets[b + 1]; ++x) {
                const uint32_t m = block_markers[x];
                acc += vf[2 * (size_t) m] - (e->has_centre ? e->centre[m] : 0.0);
            }
            base[b] = acc;
        }
        GSB_TRY(to_device_vec(frag, &e->d_local));
        {
            std::vector<double> delta(M);
            for (uint32_t m = 0; m < M; ++m) delta[m] = vf[2 * (size_t) m + 1] - vf[2 * (size_t) m];
            std::vector<uint32_t> frag5;
            build_local_tc5_fragments(delta, steps, block_offsets, block_markers, exp2v, frag5);
            GSB_TRY(to_device_vec(frag5, &e->d_local_tc5));
        }
        e->local_key = key;
    }

    // Which kernel: with several effect sets the contraction is dense enough for tcgen05 (local_tc5.cu), where
    // all the sets of a pass share one A operand; GSB200_LOCAL_PATH=mma|tc5 forces either (measurements, tests).
    static const int forced_path = [] { const char* v = getenv("GSB200_LOCAL_PATH"); return !v ? 0 : strcmp(v, "tc5") == 0 ? 2 : strcmp(v, "mma") == 0 ? 1 : 0; }();
    const bool use_tc5 = forced_path == 2 || (forced_path == 0 && n_effect_sets >= kLocalTc5MinSets);
    if (use_tc5) {
        for (uint32_t q0 = 0; q0 < n_effect_sets; q0 += kLocalTc5MaxSets) {
            const uint32_t sets = std::min<uint32_t>(kLocalTc5MaxSets, n_effect_sets - q0);
            // host output: the pass's matrices are staged on the device in slabs of individuals
            const size_t slab = device_out ? (size_t) n : std::max<size_t>(64, std::min<size_t>(n, (((size_t) 1 << 30) / (2 * (size_t) n_blocks * sizeof(double))) / sets / 64 * 64));
            const size_t slab_bytes = slab * 2 * (size_t) n_blocks * sizeof(double);
            if (!device_out) GSB_TRY(ctx->s_eval[1].ensure(slab_bytes * sets));
            for (uint32_t done = 0; done < n; done += (uint32_t) slab) {
                const uint32_t cnt = (uint32_t) std::min<size_t>(slab, n - done);
                const uint32_t* d_b[kLocalTc5MaxSets];
                const double* d_base[kLocalTc5MaxSets];
                double* d_out[kLocalTc5MaxSets];
                double scale[kLocalTc5MaxSets];
                for (uint32_t q = 0; q < sets; ++q) {
                    DeviceEffects* e = effs[q0 + q];
                    d_b[q] = e->d_local_tc5;
                    d_base[q] = (const double*) ((const char*) e->d_local + frag_bytes);
                    scale[q] = e->local_scale;
                    d_out[q] = device_out ? outs[q0 + q] : (double*) ((char*) ctx->s_eval[1].ptr + slab_bytes * q);
                }
                GSB_TRY(local_tc5_launch(ctx, cnt, d_rows + done, n_blocks, steps, block_first, sets, d_b, d_base, d_out, scale));
                if (!device_out)
                    for (uint32_t q = 0; q < sets; ++q)
                        GSB_CUDA_TRY(cudaMemcpyAsync(outs[q0 + q] + (size_t) done * 2 * n_blocks, d_out[q],
                                                     (size_t) cnt * 2 * n_blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                int flag = 0;
                GSB_CUDA_TRY(cudaMemcpyAsync(&flag, ctx->d_flag + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                if (flag) {
                    GSB_CUDA_TRY(cudaMemsetAsync(ctx->d_flag + 2, 0, sizeof(int), ctx->stream));
                    GSB_REQUIRE(false, GSB200_ERR_CUDA, "%s (tcgen05 path): a pipeline wait timed out", what);
                }
            }
        }
        return GSB200_OK;
    }

    const size_t slab_rows = std::max<size_t>(8, std::min<size_t>(n, (((size_t) 1 << 30) / (2 * (size_t) n_blocks * sizeof(double))) / 8 * 8));
    for (uint32_t q0 = 0; q0 < n_effect_sets; q0 += kLocalMaxSets) {
        const uint32_t sets = std::min<uint32_t>(kLocalMaxSets, n_effect_sets - q0);
        LocalSetParams sp;
        memset(&sp, 0, sizeof sp);
        for (uint32_t q = 0; q < sets; ++q) {
            DeviceEffects* e = effs[q0 + q];
            sp.bfrag[q] = (const uint2*) e->d_local;
            sp.base[q] = (const double*) ((const char*) e->d_local + frag_bytes);
            sp.scale[q] = e->local_scale;
        }
        // individuals in slabs: the device output of `sets` matrices stays bounded
        const size_t slab = device_out ? (size_t) n : std::max<size_t>(8, slab_rows / sets / 8 * 8);
        const size_t slab_bytes = slab * 2 * (size_t) n_blocks * sizeof(double);
        if (!device_out) GSB_TRY(ctx->s_eval[1].ensure(slab_bytes * sets));
        for (uint32_t done = 0; done < n; done += (uint32_t) slab) {
            const uint32_t cnt = (uint32_t) std::min<size_t>(slab, n - done);
            for (uint32_t q = 0; q < sets; ++q) {
                sp.out[q] = device_out ? outs[q0 + q] : (double*) ((char*) ctx->s_eval[1].ptr + slab_bytes * q);
                GSB_CUDA_TRY(cudaMemsetAsync(sp.out[q], 0, (size_t) cnt * 2 * n_blocks * sizeof(double), ctx->stream));   // empty blocks read 0
            }
            const uint64_t group_blocks = ((cnt + 15) / 16 + kLocalWarps - 1) / kLocalWarps;
            GSB_REQUIRE(group_blocks * n_ranges < (1ull << 31), GSB200_ERR_ARG, "%s: too many individuals x blocks for one launch", what);
            const unsigned grid = (unsigned) (group_blocks * n_ranges);
#define GSB_LOCAL_LAUNCH(S) local_gebv_digits_kernel<S><<<grid, kLocalWarps * 32, 0, ctx->stream>>>( \
                ctx->d_store, ctx->row_words(), ctx->plane_words, d_rows + done, cnt, n_blocks, d_steps, d_ranges, n_ranges, sp)
            switch (sets) {
                case 1: GSB_LOCAL_LAUNCH(1); break;
                case 2: GSB_LOCAL_LAUNCH(2); break;
                case 3: GSB_LOCAL_LAUNCH(3); break;
                default: GSB_LOCAL_LAUNCH(4); break;
            }
#undef GSB_LOCAL_LAUNCH
            ++ctx->launches;
            GSB_CUDA_TRY(cudaGetLastError());
            if (!device_out)
                for (uint32_t q = 0; q < sets; ++q)
                    GSB_CUDA_TRY(cudaMemcpyAsync(outs[q0 + q] + (size_t) done * 2 * n_blocks, sp.out[q],
                                                 (size_t) cnt * 2 * n_blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
    }
    return GSB200_OK;
}

int gsb200_local_gebv_multi(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                            const uint32_t* block_offsets, const uint32_t* block_markers,
                            uint32_t n_effect_sets, const uint32_t* eff_indexes, double* const* outs) {
    return local_gebv_multi_impl(ctx, n, rows, n_blocks, block_offsets, block_markers, n_effect_sets, eff_indexes, outs, false);
}

int gsb200_local_gebv_multi_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                                const uint32_t* block_offsets, const uint32_t* block_markers,
                                uint32_t n_effect_sets, const uint32_t* eff_indexes, double* const* d_outs) {
    return local_gebv_multi_impl(ctx, n, rows, n_blocks, block_offsets, block_markers, n_effect_sets, eff_indexes, d_outs, true);
}

int gsb200_local_gebv(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t n_blocks,
                      const uint32_t* block_offsets, const uint32_t* block_markers,
                      uint32_t eff_index, double* out) {
    double* outs[1] = { out };
    return local_gebv_multi_impl(ctx, n, rows, n_blocks, block_offsets, block_markers, 1, &eff_index, outs, false);
}

int gsb200_allele_counts(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, char allele, int out_kind, void* out) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = "gsb200_allele_counts";
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(out != nullptr, GSB200_ERR_ARG, "%s: null output", what);
    GSB_REQUIRE(out_kind == GSB200_COUNTS_F64 || out_kind == GSB200_COUNTS_I32 || out_kind == GSB200_COUNTS_U8,
                GSB200_ERR_ARG, "%s: unknown output kind %d", what, out_kind);
    GSB_TRY(check_rows(ctx, n, rows, what));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    GSB_TRY(sync_dictionary(ctx));
    uint32_t* d_rows;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], rows, n, &d_rows));
    const size_t elt = out_kind == GSB200_COUNTS_F64 ? 8 : (out_kind == GSB200_COUNTS_I32 ? 4 : 1);
    const size_t row_out = (size_t) ctx->n_markers * elt;
    const uint32_t slab = (uint32_t) std::max<size_t>(1, std::min<size_t>(n, ((size_t) 1 << 30) / row_out));
    GSB_TRY(ctx->s_eval[1].ensure((size_t) slab * row_out));
    for (uint32_t done = 0; done < n; done += slab) {
        const uint32_t cnt = std::min(slab, n - done);
        const uint64_t threads = (uint64_t) cnt * ctx->n_markers;
        const unsigned grid = (unsigned) ((threads + 255) / 256);
        if (out_kind == GSB200_COUNTS_F64)
            allele_count_kernel<double><<<grid, 256, 0, ctx->stream>>>(ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes,
                ctx->n_markers, d_rows + done, cnt, ctx->d_dict_sym, (uint8_t) allele, (double*) ctx->s_eval[1].ptr);
        else if (out_kind == GSB200_COUNTS_I32)
            allele_count_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes,
                ctx->n_markers, d_rows + done, cnt, ctx->d_dict_sym, (uint8_t) allele, (int32_t*) ctx->s_eval[1].ptr);
        else
            allele_count_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes,
                ctx->n_markers, d_rows + done, cnt, ctx->d_dict_sym, (uint8_t) allele, (uint8_t*) ctx->s_eval[1].ptr);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
        GSB_CUDA_TRY(cudaMemcpyAsync((char*) out + (size_t) done * row_out, ctx->s_eval[1].ptr, (size_t) cnt * row_out,
                                     cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return GSB200_OK;
}

int gsb200_allele_presence(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint8_t* present) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = "gsb200_allele_presence";
    GSB_REQUIRE(present != nullptr, GSB200_ERR_ARG, "%s: null output", what);
    const uint32_t M = ctx->n_markers, planes = ctx->planes, codes = 1u << planes;
    memset(present, 0, (size_t) M << planes);
    if (n == 0) return GSB200_OK;
    GSB_TRY(check_rows(ctx, n, rows, what));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t* d_rows;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], rows, n, &d_rows));
    const size_t bytes = (size_t) codes * ctx->plane_words * sizeof(uint32_t);
    GSB_TRY(ctx->s_eval[1].ensure(bytes));
    GSB_CUDA_TRY(cudaMemsetAsync(ctx->s_eval[1].ptr, 0, bytes, ctx->stream));
    const uint32_t rows_per_slab = 256;
    dim3 grid((ctx->plane_words + 127) / 128, (n + rows_per_slab - 1) / rows_per_slab);
    presence_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->d_store, ctx->row_words(), ctx->plane_words, planes, d_rows, n, rows_per_slab,
                                                   (uint32_t*) ctx->s_eval[1].ptr);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> masks((size_t) codes * ctx->plane_words);
    GSB_CUDA_TRY(cudaMemcpyAsync(masks.data(), ctx->s_eval[1].ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (uint32_t c = 0; c < codes; ++c)
        for (uint32_t m = 0; m < M; ++m)
            present[((size_t) m << planes) + c] = (masks[(size_t) c * ctx->plane_words + (m >> 5)] >> (m & 31)) & 1u;
    return GSB200_OK;
}

int gsb200_top_n(gsb200_ctx* ctx, uint32_t n, const double* values, uint32_t top_n, int low_is_best, uint32_t* out_positions) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = "gsb200_top_n";
    if (n == 0 || top_n == 0) return GSB200_OK;
    GSB_REQUIRE(values && out_positions, GSB200_ERR_ARG, "%s: null array", what);
    GSB_REQUIRE(top_n <= n, GSB200_ERR_ARG, "%s: top_n %u exceeds n %u", what, top_n, n);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    // s_eval[3]: values | selected positions
    GSB_TRY(ctx->s_eval[3].ensure((size_t) n * 8 + (size_t) top_n * 4));
    GSB_TRY(ctx->s_select.ensure(select_scratch_bytes(n)));
    double* d_val = (double*) ctx->s_eval[3].ptr;
    uint32_t* d_sel = (uint32_t*) ((char*) ctx->s_eval[3].ptr + (size_t) n * 8);
    GSB_CUDA_TRY(cudaMemcpyAsync(d_val, values, (size_t) n * 8, cudaMemcpyHostToDevice, ctx->stream));
    GSB_TRY(select_top_n(ctx, d_val, n, top_n, low_is_best, ctx->s_select.ptr, d_sel, 0, n));
    GSB_CUDA_TRY(cudaMemcpyAsync(out_positions, d_sel, (size_t) top_n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

}  // extern "C"
