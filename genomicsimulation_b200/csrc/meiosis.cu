// meiosis.cu — batched gamete generation on the bit-packed store (K1-K4).
//
// Replaces gsc_generate_gamete (core.c:7850-7928), gsc_generate_doubled_haploid
// (core.c:7949-8033), gsc_generate_clone (core.c:8047-8055) and the inner double loop of
// gsc_scaffold_make_new_genotypes (core.c:8301-8330) for every offspring of a batch at once.
//
// The reference walks the markers of a chromosome and flips the strand it copies from while
// `dists[i] > crossover[k]` (core.c:7902-7903).  Equivalently (see context.cu): with d[] the
// running maximum of dists, the strand at marker i is
//        start_strand  XOR  parity( #{ j : d[i] > x_j } )
// so every crossover j, independently of the others and of their order, flips the strand on
// the marker range [ub_j, end of chromosome), ub_j = first i with d[i] > x_j, and a start
// strand of 1 flips the whole chromosome.  A gamete is therefore described by a small
// multiset of "toggles" (marker positions where the strand mask flips from there to the end
// of the genome; every range contributes one toggle at its start and one at its end), and
// the strand mask of a 128-marker piece is: parity of the toggles before it, XOR a partial
// mask per toggle inside it.  No sorting is needed — XOR commutes.
//
// Fast kernel (maps whose linkage groups are contiguous, disjoint and cover every marker):
// one warp per gamete.  The warp draws (or reads from the replay tape) the crossover counts
// lane-per-chromosome, the positions lane-per-crossover, locates each crossover with a
// bucketed upper_bound on the reference's own fp64 dists, bins the toggles by 128-marker
// piece with shared-memory atomics, prefix-sums the bins, and then streams the plane with
// 128-bit loads and stores, reading only the parental strand(s) a piece actually inherits.
//
// General kernel (REORDER linkage groups, maps that do not cover every marker, maps with so
// many linkage groups that the toggle list does not fit shared memory — e.g. the unlinked
// map of core.c:6128-6152): one thread per 32 consecutive map positions, gathers bit by bit.
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cub/cub.cuh>

using namespace gsb;

namespace {

constexpr uint32_t kMaxNativeGenerations = 2048;   // generation field of the Philox counter
constexpr uint32_t kMaxNativeCrossovers = (1u << 20) - 2;
constexpr double   kMaxNativeLambda = 500.0;

struct MapView {
    const ChrInfo* chr;
    const double* dists;
    const uint32_t* bucket;
    const uint32_t* marker_index;
    const uint32_t* chunk_chr;
    const uint32_t* chunk_start;
    uint32_t n_chr;
    uint32_t n_chunks;
};

struct DrawView {
    uint32_t mode;             // GSB200_DRAWS_*
    uint32_t key0, key1;       // PHILOX
    uint64_t first_unit;
    uint32_t generation;
    // TAPE layout (host tape uploaded, or native draws materialised by export kernels)
    const uint32_t* tape_k;
    const uint8_t*  tape_s;
    const uint64_t* tape_posoff;
    const double*   tape_pos;
    uint64_t unit_stride;      // tape entries per unit
    uint64_t base_offset;      // entries before this generation's gametes within a unit
};

constexpr long long kPoolWaitCycles = 40000000000ll;        // ~20 s: a peer that never delivers (shard.cu has the same)

struct MeiosisArgs {
    const uint32_t* src_base;  // rows the gametes are made from
    const uint32_t* src_rows0; // per unit: row of the parent of gamete 0 (null: unit index)
    const uint32_t* src_rows1; // ... of gamete 1
    uint32_t* dst_base;
    const uint32_t* dst_rows;  // per unit (null: unit index)
    uint64_t row_words;
    uint32_t plane_words, planes, n_markers;
    uint32_t n_units;
    uint32_t gametes_per_unit; // 2: cross / selfing generation, 1: doubled haploid
    uint32_t doubled;          // write the gamete to both strands
    uint32_t fix_chunked;      // fast kernel: apply the toggles' fix-ups chunk by chunk, while the lines are still in L2 (see E1/E2)
    MapView map[2];
    DrawView draws;
    // fast-kernel plan geometry
    uint32_t n_u;              // 128-marker pieces (uint4) per plane
    uint32_t parity_words;     // ceil(n_u / 32)
    uint32_t chr_words;        // ceil(max n_chr / 32)
    uint32_t xcap, tcap;       // crossovers / toggles one gamete's plan can hold
    uint32_t warp_smem_words;
    int* flag;
    // fused GEBV (splice_flat_kernel<true>): see parent_tables_kernel
    const long long* bv_tab;   // [slot][(plane_words * 32 >> bv_shift) + 2]
    uint32_t bv_shift;         // one table entry per 2^bv_shift markers (3 or 5)
    const uint32_t* bv_slot;   // row -> slot
    const long long* bv_q;     // [plane_words * 32]: fixed-point delta per marker
    unsigned long long* bv_acc;// per unit: sum over its gametes, two's complement
    const uint32_t* order;     // fast kernel: the gametes in processing order (null: as numbered); see pool_random_cross
    PoolWait pw;               // fast kernel: flags null unless the source rows are a pool that is still arriving
};

// ------------------------------------------------------------------ native draw stream

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// uniform in the OPEN interval (0,1): (v + 1/2) / 2^52 with v the top 52 bits, so the extremes are
// 2^-53 and 1 - 2^-53, both exact in fp64 (with 53 bits, (2^53 - 1) + 1/2 would round up to 2^53 and
// give exactly 1.0 once in 2^53 draws)
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    const uint64_t v = (((uint64_t) hi << 32) | lo) >> 12;
    return ((double) v + 0.5) * (1.0 / 4503599627370496.0);
}

__device__ __forceinline__ uint4 draw_block(const DrawView& d, uint64_t unit, uint32_t gamete, uint32_t chr, uint32_t index) {
    const uint4 ctr = make_uint4((uint32_t) unit, (uint32_t) (unit >> 32), chr,
                                 (index << 12) | (d.generation << 1) | gamete);
    return philox4x32_10(ctr, d.key0, d.key1);
}

__constant__ double kInverse[32] = {
    0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13,
    1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25,
    1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31 };

// Poisson(lambda) by inversion of one uniform; p0 = exp(-lambda).
__device__ __forceinline__ uint32_t poisson_inverse(double u, double lambda, double p0) {
    if (!(lambda > 0.0)) return 0;                 // Rf_rpois(0) == 0 (unlinked maps)
    double p = p0, cdf = p0;
    uint32_t k = 0;
    while (u > cdf && k < kMaxNativeCrossovers) {
        ++k;
        p *= k < 32 ? lambda * kInverse[k] : lambda / (double) k;
        cdf += p;
        if (p < 1e-300 && k > lambda) break;       // tail exhausted in fp64
    }
    return k;
}

// (count, start strand) of one gamete-chromosome: the Rf_rpois and `which` draws of
// core.c:7873 and 7896.
__device__ __forceinline__ void draw_header(const DrawView& d, const ChrInfo& ci, uint64_t unit, uint32_t gamete,
                                            uint32_t chr, uint64_t entry, uint32_t& k, uint32_t& strand) {
    if (d.mode == GSB200_DRAWS_PHILOX) {
        const uint4 r = draw_block(d, unit, gamete, chr, 0);
        // (a 16-entry CDF table searched by bisection was tried: its dependent loads cost more than this short loop)
        k = poisson_inverse(u01(r.x, r.y), ci.lambda, ci.exp_neg_lambda);
        strand = r.z >> 31;
    } else {
        k = d.tape_k[entry];
        strand = d.tape_s[entry] ? 1u : 0u;
    }
}

// position of crossover j of one gamete-chromosome: the unif_rand of core.c:7889.
__device__ __forceinline__ double draw_position(const DrawView& d, uint64_t unit, uint32_t gamete, uint32_t chr,
                                                uint64_t entry, uint32_t j) {
    if (d.mode == GSB200_DRAWS_PHILOX) {
        const uint4 r = draw_block(d, unit, gamete, chr, j + 1);
        return u01(r.x, r.y);
    }
    return d.tape_pos[d.tape_posoff[entry] + j];
}

// first i in [0, len) with d[i] > x, else len — the marker at which the reference's walk
// passes crossover x (core.c:7902-7903).
__device__ __forceinline__ uint32_t first_marker_past(const ChrInfo& ci, const MapView& mv, double x) {
    if (!(x == x)) return ci.len;                  // NaN never compares below a dist
    const double* d = mv.dists + ci.pos_off;
    uint32_t lo = 0, hi = ci.len;
    if (x >= 0.0 && x < 1.0) {
        const uint32_t nb = 1u << ci.bucket_shift;
        const uint32_t b = min((uint32_t) (x * (double) nb), nb - 1);
        lo = mv.bucket[ci.bucket_off + b];
        hi = mv.bucket[ci.bucket_off + b + 1];
    }
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (d[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ uint64_t entry_index(const DrawView& d, uint64_t unit_local, uint32_t gamete, uint32_t n_chr0, uint32_t chr) {
    return unit_local * d.unit_stride + d.base_offset + (gamete ? n_chr0 : 0) + chr;
}

// ------------------------------------------------------------------ fast kernel

constexpr int kUnroll = 4;
constexpr uint32_t kNoToggle = 0xffffffffu;

// x << n with n in [0, 32]: PTX shl.b32 clamps shift amounts above 31, giving 0 at n = 32
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, int n) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
    return r;
}

__device__ __forceinline__ uint4 ldg128(const uint32_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Append the toggles the lanes hold (kNoToggle = none) to the warp's list, and flip the parity
// bit of the 128-marker piece each one falls in.
__device__ __forceinline__ void emit_toggles(uint32_t t, uint32_t lane, const MeiosisArgs& a, uint32_t* piece_parity,
                                             uint32_t* list, uint32_t& n_list) {
    const bool has = t < a.n_markers;              // nothing lies beyond the last marker
    const uint32_t ballot = __ballot_sync(0xffffffffu, has);
    if (has) {
        const uint32_t idx = n_list + __popc(ballot & ((1u << lane) - 1u));
        if (idx < a.tcap) list[idx] = t;
        atomicXor(&piece_parity[t >> 12], 1u << ((t >> 7) & 31));
    }
    n_list += __popc(ballot);
}

// One warp per gamete.  Shared memory per warp:
//   piece_parity[n_u/32 + 1]  bit u: parity of the toggles inside piece u, then (after the scan)
//                             parity of the toggles BEFORE piece u = the strand piece u starts on
//   chr_flip[n_chr/32 + 1]    bit c: parity of the crossovers that take effect in chromosome c
//   chr_start[n_chr/32 + 1]   bit c: start strand of chromosome c
//   list[tcap]                the toggles (marker positions), unordered
//   xinfo[xcap]               crossover -> (chromosome, index within chromosome)
//
// kBv: also the gamete's contribution to its offspring's GEBV, WITHOUT reading the gamete back.
// With Q_m the fixed-point effect difference of marker m and F_s(x) = sum_{m < x} Q_m * h_s(m) for
// the parent's strand s, a gamete that runs on strand s_k between toggles t_k and t_{k+1} is worth
//   sum_k F_{s_k}(t_{k+1}) - F_{s_k}(t_k)  =  F_{s_last}(M) + sum_toggles +-(F_0 - F_1)(t),
// + when the toggle leaves strand 0, - when it leaves strand 1.  (F_0 - F_1) at every 8- or 32-marker boundary
// and F_0(M) come from the parent's prefix table (parent_tables_kernel); the < 32 markers between
// the word boundary and the toggle are summed here from the parent's words.  The strand a toggle
// leaves = the strand its 128-marker piece starts on, flipped once per earlier toggle in the
// same piece (rare: found with two bitsets, ranked by scanning the list only then).
template <bool kBv, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) splice_flat_kernel(const MeiosisArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint2* xinfo = reinterpret_cast<uint2*>(smem + (size_t) warp * a.warp_smem_words);
    uint32_t* list = reinterpret_cast<uint32_t*>(xinfo + a.xcap);
    uint32_t* piece_parity = list + a.tcap;
    uint32_t* chr_flip = piece_parity + a.parity_words;
    uint32_t* chr_start = chr_flip + a.chr_words;
    uint32_t* piece_seen = chr_start + a.chr_words;          // kBv only: pieces holding >= 1 / >= 2 toggles
    uint32_t* piece_multi = piece_seen + a.parity_words;

    const uint32_t n_gametes = a.n_units * a.gametes_per_unit;      // host keeps this below 2^32
    uint32_t owners_seen = 0;                                        // pool still arriving: the ranks whose rows this warp has seen
    const uint32_t my_bound = (a.pw.flags && lane < a.pw.world) ? a.pw.bounds[lane] : 0u;
    for (uint32_t gi = blockIdx.x * warps + warp; gi < n_gametes; gi += gridDim.x * warps) {
        const uint32_t g = a.order ? a.order[gi] : gi;           // gametes sorted by parent: neighbouring warps share their source rows
        const uint32_t unit_local = a.gametes_per_unit == 2 ? g >> 1 : g;
        const uint32_t gam = a.gametes_per_unit == 2 ? g & 1u : 0u;
        const uint64_t unit = a.draws.first_unit + unit_local;
        const MapView& mv = a.map[gam];
        const uint64_t entry0 = (uint64_t) unit_local * a.draws.unit_stride + a.draws.base_offset + (gam ? a.map[0].n_chr : 0);

        for (uint32_t i = lane; i < a.parity_words + 2 * a.chr_words; i += 32) piece_parity[i] = 0;
        __syncwarp();

        // B. lane per chromosome: crossover count and start strand
        uint32_t n_x = 0;
        for (uint32_t c0 = 0; c0 < mv.n_chr; c0 += 32) {
            const uint32_t c = c0 + lane;
            uint32_t k = 0, strand = 0;
            if (c < mv.n_chr) {
                const ChrInfo ci = mv.chr[c];
                draw_header(a.draws, ci, unit, gam, c, entry0 + c, k, strand);
                if (ci.len == 0) strand = 0;
                if (ci.len == 0 || !ci.searchable) k = 0;   // such positions are drawn but can never be passed
            }
            const uint32_t sb = __ballot_sync(0xffffffffu, strand != 0);
            if (lane == 0) chr_start[c0 >> 5] = sb;
            uint32_t incl = k;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int) lane >= o) incl += v;
            }
            const uint32_t off = n_x + incl - k;
            for (uint32_t j = 0; j < k; ++j)
                if (off + j < a.xcap) xinfo[off + j] = make_uint2(c, j);
            n_x += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        if (n_x > a.xcap) {                        // host re-runs the batch on the general kernel
            if (lane == 0) atomicExch(a.flag, 1);
            continue;
        }

        // C. lane per crossover: position -> the marker from which the strand is flipped
        uint32_t n_list = 0;
        for (uint32_t i0 = 0; i0 < n_x; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t t = kNoToggle;
            if (i < n_x) {
                const uint2 xi = xinfo[i];
                const ChrInfo ci = mv.chr[xi.x];
                const double x = draw_position(a.draws, unit, gam, xi.x, entry0 + xi.x, xi.y);
                const uint32_t ub = first_marker_past(ci, mv, x);
                if (ub < ci.len) {
                    t = ci.first + ub;
                    atomicXor(&chr_flip[xi.x >> 5], 1u << (xi.x & 31));
                }
            }
            emit_toggles(t, lane, a, piece_parity, list, n_list);
        }
        __syncwarp();
        // chromosome boundaries: enter on the start strand, leave back on strand 0.  Where one chromosome
        // ends on strand 1 exactly where the next starts on strand 1 the two toggles cancel and are not made.
        {
            uint32_t carry_e = 0, carry_pos = 0;              // the previous group's last chromosome
            for (uint32_t c0 = 0; c0 < mv.n_chr; c0 += 32) {
                const uint32_t c = c0 + lane;
                uint32_t s_c = 0, e_c = 0, first = 0, end = 0;
                if (c < mv.n_chr) {
                    s_c = (chr_start[c0 >> 5] >> lane) & 1u;
                    e_c = s_c ^ ((chr_flip[c0 >> 5] >> lane) & 1u);
                    const ChrInfo ci = mv.chr[c];
                    first = ci.first;
                    end = ci.first + ci.len;
                }
                uint32_t e_prev = __shfl_up_sync(0xffffffffu, e_c, 1), end_prev = __shfl_up_sync(0xffffffffu, end, 1);
                if (lane == 0) { e_prev = carry_e; end_prev = carry_pos; }
                uint32_t t_in = kNoToggle, t_out = kNoToggle;    // t_out: the PREVIOUS chromosome's way out
                if (c < mv.n_chr) {
                    if (end_prev == first) { if (s_c ^ e_prev) t_in = first; }
                    else { if (s_c) t_in = first; if (e_prev) t_out = end_prev; }
                }
                emit_toggles(t_in, lane, a, piece_parity, list, n_list);
                emit_toggles(t_out, lane, a, piece_parity, list, n_list);
                const uint32_t last = min(31u, mv.n_chr - 1 - c0);
                carry_e = __shfl_sync(0xffffffffu, e_c, last);
                carry_pos = __shfl_sync(0xffffffffu, end, last);
            }
            emit_toggles(lane == 0 && carry_e ? carry_pos : kNoToggle, lane, a, piece_parity, list, n_list);
        }
        __syncwarp();
        if (n_list > a.tcap) {
            if (lane == 0) atomicExch(a.flag, 1);
            continue;
        }

        // D. exclusive prefix XOR over the piece parities: bit u = strand piece u starts on
        {
            uint32_t carry = 0;
            for (uint32_t w0 = 0; w0 < a.parity_words; w0 += 32) {
                const uint32_t w = w0 + lane;
                uint32_t x = w < a.parity_words ? piece_parity[w] : 0u;
                uint32_t incl = x;
                incl ^= incl << 1; incl ^= incl << 2; incl ^= incl << 4; incl ^= incl << 8; incl ^= incl << 16;
                uint32_t word_par = incl >> 31, pre = word_par;      // parity of the whole word
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, pre, o);
                    if ((int) lane >= o) pre ^= v;
                }
                const uint32_t before = carry ^ pre ^ word_par;      // parity of all earlier words
                if (w < a.parity_words) piece_parity[w] = (incl << 1) ^ (before ? 0xffffffffu : 0u);
                carry ^= __shfl_sync(0xffffffffu, pre, 31);
            }
        }
        __syncwarp();

        const uint32_t src_row = gam ? (a.src_rows1 ? a.src_rows1[unit_local] : unit_local)
                                     : (a.src_rows0 ? a.src_rows0[unit_local] : unit_local);
        const uint32_t dst_row = a.dst_rows ? a.dst_rows[unit_local] : unit_local;
        const uint32_t* src = a.src_base + (uint64_t) src_row * a.row_words;
        uint32_t* dst = a.dst_base + (uint64_t) dst_row * a.row_words;
        const uint32_t hap_stride = a.planes * a.plane_words;

        // The pool may still be arriving from the peers of a sharded generation: wait for the OWNER of this parent's slot
        // only (everything above — draws, crossover plan — needed no parent data).  The flag is read with acquire semantics
        // by lane 0; the other lanes pass the warp barrier after it, so no row load below is issued before the flag was seen.
        // (A __threadfence() here instead cost 15 % of the kernel: it also waits for the warp's own stores of the gamete before.)
        if (a.pw.flags) {
            // lane r holds the first slot of rank r (read once per warp): the owner is the last rank whose first slot is <= src_row
            const uint32_t owner = (uint32_t) __popc(__ballot_sync(0xffffffffu, lane > 0 && lane < a.pw.world && src_row >= my_bound));
            if (!((owners_seen >> owner) & 1u)) {                  // (warp-uniform: a rank's rows, once here, stay)
                if (lane == 0 && owner != a.pw.rank) {
                    const long long t0 = clock64();
                    for (;;) {
                        const uint32_t have = a.pw.flags[owner];          // volatile: read past L1
                        if (have >= a.pw.epoch) break;
                        if (clock64() - t0 > kPoolWaitCycles) { atomicExch(a.pw.err, 1); break; }
                        __nanosleep(200);
                    }
                }
                __syncwarp();
                owners_seen |= 1u << owner;
            }
        }

        // E1. every piece, whole, from the strand it starts on: plain 128-bit copies.  Full groups of
        //     32 * kUnroll pieces run without bounds checks; 32-bit piece offsets from two bases.
        const uint32_t hap_pieces = hap_stride >> 2;
        const uint32_t full = a.n_u & ~(32u * kUnroll - 1u);
        // With rows too long for everything the resident warps have in flight to stay in L2 (M = 100 k: 25 - 37 KB per gamete x
        // 5 920 warps > 126 MB), fix-ups applied after the WHOLE gamete find their lines evicted again: every toggle then re-reads
        // a sector of each strand from DRAM and its atomics read and write back the output's (ncu: 25 KB read per doubled haploid
        // where 12.5 KB are needed).  fix_chunked applies them after every 2 KB chunk instead (one plane; not the GEBV variant).
        const bool chunked = !kBv && a.fix_chunked && a.planes == 1;
        for (uint32_t p = 0; p < a.planes && !chunked; ++p) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src + p * a.plane_words);
            uint4* o4 = reinterpret_cast<uint4*>(dst + (gam * a.planes + p) * a.plane_words);
            uint32_t u = lane;
            for (; u < full; u += 32 * kUnroll) {
                uint4 v[kUnroll];
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    const uint32_t on1 = (piece_parity[(u >> 5) + q] >> lane) & 1u;
                    v[q] = __ldg(s4 + (u + q * 32 + on1 * hap_pieces));
                }
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    o4[u + q * 32] = v[q];
                    if (a.doubled) o4[u + q * 32 + hap_pieces] = v[q];
                }
            }
            for (; u < a.n_u; u += 32) {
                const uint32_t on1 = (piece_parity[u >> 5] >> lane) & 1u;
                const uint4 v = __ldg(s4 + (u + on1 * hap_pieces));
                o4[u] = v;
                if (a.doubled) o4[u + hap_pieces] = v;
            }
        }
        __syncwarp();      // orders the plain stores before the fix-ups below (same warp)

        // E2. lane per toggle: flipping the strand from bit o of its piece on means XOR-ing
        //     (strand0 ^ strand1) & (ones from bit o) into what E1 wrote; XOR fix-ups of several
        //     toggles in one piece commute, so they are applied with 64-bit atomic XORs.
        long long bv = 0;
        const long long* bv_tab = nullptr;
        uint32_t* multi_keys = reinterpret_cast<uint32_t*>(xinfo);       // the crossover list is dead by now
        uint32_t n_multi = 0;
        if (kBv) {
            uint32_t slot = a.bv_slot[src_row];
            if (slot == 0xffffffffu) {                       // parent outside the pool: the GEBV is void
                if (lane == 0) atomicExch(a.flag, 4);
                slot = 0;
            }
            bv_tab = a.bv_tab + (size_t) slot * (((a.plane_words * 32u) >> a.bv_shift) + 2);
            // pieces holding two or more toggles, and the (few) toggles inside them as keys (position, index)
            for (uint32_t i = lane; i < 2 * a.parity_words; i += 32) piece_seen[i] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < n_list; i += 32) {
                const uint32_t u = list[i] >> 7, bit = 1u << (u & 31);
                if (atomicOr(&piece_seen[u >> 5], bit) & bit) atomicOr(&piece_multi[u >> 5], bit);
            }
            __syncwarp();
            for (uint32_t i0 = 0; i0 < n_list; i0 += 32) {
                const uint32_t i = i0 + lane;
                const uint32_t t = i < n_list ? list[i] : 0u;
                const bool in_multi = i < n_list && ((piece_multi[t >> 12] >> ((t >> 7) & 31)) & 1u);
                const uint32_t ballot = __ballot_sync(0xffffffffu, in_multi);
                if (in_multi) {
                    const uint32_t k = n_multi + __popc(ballot & ((1u << lane) - 1u));
                    if (k < 2 * a.xcap) multi_keys[k] = (t << 7) | (i & 127u);
                }
                n_multi += __popc(ballot);
            }
            __syncwarp();
            if (n_multi > 2 * a.xcap || n_list > 128) {      // cannot rank them here: the batch is redone
                if (lane == 0) atomicExch(a.flag, 1);
                n_multi = 0;
            }
        }
        auto fix_pieces = [&](const uint32_t u_lo, const uint32_t u_hi) {
        for (uint32_t i0 = 0; i0 < n_list; i0 += 32) {
            const uint32_t i = i0 + lane;
            if (i >= n_list) continue;
            const uint32_t t = list[i], u = t >> 7;
            if (u < u_lo || u >= u_hi) continue;
            const int o = (int) (t & 127u);
            const uint32_t m0 = shl_clamp(0xffffffffu, o), m1 = shl_clamp(0xffffffffu, max(o - 32, 0));
            const uint32_t m2 = shl_clamp(0xffffffffu, max(o - 64, 0)), m3 = shl_clamp(0xffffffffu, max(o - 96, 0));
            for (uint32_t p = 0; p < a.planes; ++p) {
                const uint32_t* s0 = src + p * a.plane_words + 4 * u;
                const uint4 h0 = ldg128(s0), h1 = ldg128(s0 + hap_stride);
                if (kBv) {
                    // strand this toggle leaves: the one its piece starts on, flipped per earlier toggle in the piece
                    uint32_t leaves = (piece_parity[u >> 5] >> (u & 31)) & 1u;
                    if ((piece_multi[u >> 5] >> (u & 31)) & 1u) {
                        const uint32_t key = (t << 7) | (i & 127u);
                        for (uint32_t k = 0; k < n_multi; ++k) {
                            const uint32_t kj = multi_keys[k];
                            leaves ^= (uint32_t) ((kj >> 14) == u && kj < key);
                        }
                    }
                    // the prefix table entry below the toggle, then the markers between that boundary and the toggle
                    const uint32_t span = (1u << a.bv_shift) - 1u;
                    const uint32_t wq = (uint32_t) o >> 5, low = ((1u << (o & span)) - 1u) << (o & 31u & ~span);
                    const uint32_t x0 = wq == 0 ? h0.x : wq == 1 ? h0.y : wq == 2 ? h0.z : h0.w;
                    const uint32_t x1 = wq == 0 ? h1.x : wq == 1 ? h1.y : wq == 2 ? h1.z : h1.w;
                    long long gd = bv_tab[t >> a.bv_shift];
                    // markers below the toggle in its word where the strands differ: four table reads in
                    // flight at a time (one read per trip would cost a full cache latency per marker)
                    const long long* qw = a.bv_q + ((size_t) (t >> 5) << 5);
                    for (uint32_t m = (x0 ^ x1) & low; m;) {
                        uint32_t b[4];
                        long long v[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            b[k] = m ? (uint32_t) __ffs((int) m) - 1u : 32u;
                            m &= m - 1;                      // 0 stays 0
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) v[k] = b[k] < 32u ? qw[b[k]] : 0ll;
#pragma unroll
                        for (int k = 0; k < 4; ++k) gd += ((x0 >> (b[k] & 31u)) & 1u) ? v[k] : -v[k];
                    }
                    bv += leaves ? -gd : gd;
                }
                const unsigned long long lo = ((unsigned long long) ((h0.y ^ h1.y) & m1) << 32) | ((h0.x ^ h1.x) & m0);
                const unsigned long long hi = ((unsigned long long) ((h0.w ^ h1.w) & m3) << 32) | ((h0.z ^ h1.z) & m2);
                unsigned long long* out = reinterpret_cast<unsigned long long*>(dst + (gam * a.planes + p) * a.plane_words + 4 * u);
                if (lo) atomicXor(out, lo);
                if (hi) atomicXor(out + 1, hi);
                if (a.doubled) {
                    unsigned long long* out2 = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint32_t*>(out) + hap_stride);
                    if (lo) atomicXor(out2, lo);
                    if (hi) atomicXor(out2 + 1, hi);
                }
            }
        }
        };
        if (chunked) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* o4 = reinterpret_cast<uint4*>(dst + gam * a.plane_words);
            uint32_t u0 = 0, fixed_to = 0, groups = 0;
            for (; u0 < full; u0 += 32 * kUnroll) {
                const uint32_t u = u0 + lane;
                uint4 v[kUnroll];
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    const uint32_t on1 = (piece_parity[(u >> 5) + q] >> lane) & 1u;
                    v[q] = __ldg(s4 + (u + q * 32 + on1 * hap_pieces));
                }
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    o4[u + q * 32] = v[q];
                    if (a.doubled) o4[u + q * 32 + hap_pieces] = v[q];
                }
                if (++groups == a.fix_chunked) {          // fix_chunked = groups of 2 KB per round of fix-ups
                    __syncwarp();
                    fix_pieces(fixed_to, u0 + 32 * kUnroll);
                    fixed_to = u0 + 32 * kUnroll;
                    groups = 0;
                }
            }
            for (uint32_t u = u0 + lane; u < a.n_u; u += 32) {
                const uint32_t on1 = (piece_parity[u >> 5] >> lane) & 1u;
                const uint4 v = __ldg(s4 + (u + on1 * hap_pieces));
                o4[u] = v;
                if (a.doubled) o4[u + hap_pieces] = v;
            }
            __syncwarp();
            fix_pieces(fixed_to, a.n_u);
        } else {
            fix_pieces(0u, 0xffffffffu);
        }
        if (kBv) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bv += __shfl_xor_sync(0xffffffffu, bv, o);
            if (lane == 0) {
                // the walk starts on strand 0; an odd number of toggles leaves it on strand 1 at the end
                const uint32_t n_tab = (a.plane_words * 32u) >> a.bv_shift;
                bv += bv_tab[n_tab + 1] - ((n_list & 1u) ? bv_tab[n_tab] : 0ll);
                atomicAdd(a.bv_acc + unit_local, (unsigned long long) bv);
            }
        }
        __syncwarp();
    }
}

// Prefix tables of the pool rows, for splice_flat_kernel<true>, one entry per 8 markers while the
// tables stay small (a coarser table leaves the splice kernel a longer marker-by-marker walk from
// the boundary to the toggle: half of the fused path's extra instructions at 32), per 32 otherwise:
//   tab[slot][k]          = sum_{m < G k} Q_m * (h0(m) - h1(m)),   G = 8 or 32, k = 0 .. n_tab = 32 * plane_words / G
//   tab[slot][n_tab + 1]  = sum_m Q_m * h0(m)
// and slot_of_row[row] = slot.  First every byte's own sum: a block owns 32 marker words, keeps
// their 1024 Q values in shared memory and walks the pool rows, 8 at a time, a thread per
// (row, word).  Then one block per row turns the sums into prefixes.
constexpr uint32_t kTabWords = 32, kTabRows = 8;
__global__ void __launch_bounds__(kTabWords * kTabRows) parent_words_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ pool_rows, uint32_t n_pool, uint32_t rows_per_block, const long long* __restrict__ q,
        uint32_t per_word /* table entries per marker word: 1 or 4 */, long long* __restrict__ tab, uint32_t* __restrict__ slot_of_row) {
    __shared__ long long s_q[kTabWords * 32];
    const uint32_t w_local = threadIdx.x & (kTabWords - 1), r_local = threadIdx.x / kTabWords;
    const uint32_t w = blockIdx.x * kTabWords + w_local;
    for (uint32_t i = threadIdx.x; i < kTabWords * 32; i += kTabWords * kTabRows) {
        const size_t m = (size_t) blockIdx.x * kTabWords * 32 + i;
        s_q[i] = m < (size_t) plane_words * 32 ? q[m] : 0ll;
    }
    __syncthreads();
    const uint32_t slot_lo = blockIdx.y * rows_per_block, slot_hi = min(slot_lo + rows_per_block, n_pool);
    const long long* qw = s_q + w_local * 32;
    for (uint32_t slot = slot_lo + r_local; slot < slot_hi; slot += kTabRows) {
        const uint32_t row = pool_rows[slot];
        long long* out = tab + (size_t) slot * (per_word * plane_words + 2);
        if (w == 0) slot_of_row[row] = slot;
        long long d[4] = { 0, 0, 0, 0 }, t0 = 0;
        if (w < plane_words) {
            const uint32_t* h0 = store + (uint64_t) row * row_words;
            const uint32_t x0 = h0[w], x1 = h0[plane_words + w];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                for (uint32_t m = ((x0 | x1) >> (8 * j)) & 255u; m; m &= m - 1) {
                    const uint32_t b = 8 * j + (uint32_t) __ffs((int) m) - 1u;
                    const long long v = qw[b];
                    const uint32_t in0 = (x0 >> b) & 1u, in1 = (x1 >> b) & 1u;
                    if (in0) t0 += v;
                    d[j] += (long long) ((int) in0 - (int) in1) * v;
                }
            if (per_word == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) out[4 * w + j] = d[j];
            } else {
                out[w] = d[0] + d[1] + d[2] + d[3];
            }
        }
        // strand 0's total: one atomic per warp = per row (a warp is the 32 words of one row here)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        if (w_local == 0 && t0) atomicAdd(reinterpret_cast<unsigned long long*>(out + per_word * plane_words + 1), (unsigned long long) t0);
    }
}

__global__ void __launch_bounds__(256) parent_scan_kernel(uint32_t n_entries, long long* __restrict__ tab) {
    __shared__ long long s_warp[8];
    long long* out = tab + (size_t) blockIdx.x * (n_entries + 2);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long carry = 0;
    for (uint32_t w0 = 0; w0 < n_entries; w0 += 256) {
        const uint32_t w = w0 + threadIdx.x;
        const long long d = w < n_entries ? out[w] : 0ll;
        long long incl = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int) lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        long long before = carry, total = carry;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) { if (k < warp) before += s_warp[k]; total += s_warp[k]; }
        if (w < n_entries) out[w] = before + incl - d;
        carry = total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n_entries] = carry;
}

// bv[i] = offset + scale * (sum of the unit's gamete values)
__global__ void bv_finish_kernel(const long long* __restrict__ acc, uint32_t n, double scale, double offset, double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = offset + (double) acc[i] * scale;
}

// ------------------------------------------------------------------ general kernel

__global__ void zero_rows_kernel(uint32_t* base, const uint32_t* rows, uint32_t n, uint64_t row_words) {
    const uint64_t per_row = row_words / 4;
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) n * per_row) return;
    const uint32_t i = (uint32_t) (tid / per_row);
    const uint64_t row = rows ? rows[i] : i;
    reinterpret_cast<uint4*>(base + row * row_words)[tid % per_row] = make_uint4(0, 0, 0, 0);
}

// One thread per (unit, 32 consecutive positions of one linkage group) of gamete `gam`.
// Draws always come in tape layout here.  Output rows must have been zeroed.
__global__ void splice_general_kernel(const MeiosisArgs a, const uint32_t gam) {
    const MapView& mv = a.map[gam];
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) a.n_units * mv.n_chunks) return;
    const uint32_t unit_local = (uint32_t) (tid / mv.n_chunks), q = (uint32_t) (tid % mv.n_chunks);
    const uint32_t c = mv.chunk_chr[q], start = mv.chunk_start[q];
    const ChrInfo ci = mv.chr[c];
    const uint64_t entry = entry_index(a.draws, unit_local, gam, a.map[0].n_chr, c);
    const uint32_t k = ci.searchable ? a.draws.tape_k[entry] : 0;
    const uint32_t strand0 = a.draws.tape_s[entry] ? 1u : 0u;
    const double* pos = a.draws.tape_pos + a.draws.tape_posoff[entry];
    const uint32_t n_here = min(32u, ci.len - start);

    const uint32_t src_row = gam ? (a.src_rows1 ? a.src_rows1[unit_local] : unit_local)
                                 : (a.src_rows0 ? a.src_rows0[unit_local] : unit_local);
    const uint32_t dst_row = a.dst_rows ? a.dst_rows[unit_local] : unit_local;
    const uint32_t* src = a.src_base + (uint64_t) src_row * a.row_words;
    uint32_t* dst = a.dst_base + (uint64_t) dst_row * a.row_words;
    const uint32_t hap_stride = a.planes * a.plane_words;

    // strand of each of the positions
    uint32_t strands = 0;
    for (uint32_t i = 0; i < n_here; ++i) {
        const double d = mv.dists[ci.pos_off + start + i];
        uint32_t passed = 0;
        for (uint32_t j = 0; j < k; ++j) passed += (d > pos[j]) ? 1u : 0u;
        strands |= ((strand0 ^ passed) & 1u) << i;
    }
    for (uint32_t p = 0; p < a.planes; ++p) {
        uint32_t cur_word = 0xffffffffu, acc = 0;
        for (uint32_t i = 0; i < n_here; ++i) {
            const uint32_t m = mv.marker_index[ci.pos_off + start + i];
            const uint32_t s = (strands >> i) & 1u;
            const uint32_t bit = (src[(uint64_t) (s * a.planes + p) * a.plane_words + (m >> 5)] >> (m & 31)) & 1u;
            if ((m >> 5) != cur_word) {
                if (acc) {
                    uint32_t* out = dst + (uint64_t) (gam * a.planes + p) * a.plane_words + cur_word;
                    atomicOr(out, acc);
                    if (a.doubled) atomicOr(out + hap_stride, acc);
                }
                cur_word = m >> 5;
                acc = 0;
            }
            acc |= bit << (m & 31);
        }
        if (acc) {
            uint32_t* out = dst + (uint64_t) (gam * a.planes + p) * a.plane_words + cur_word;
            atomicOr(out, acc);
            if (a.doubled) atomicOr(out + hap_stride, acc);
        }
    }
}

// Markers the gamete's map does not cover are never written by the reference and keep the zeroed
// slot's '\0' (core.c:75): write the code that decodes to '\0' there (rows were zeroed first).
__global__ void fill_uncovered_kernel(const MeiosisArgs a, const uint32_t gam, const uint32_t* __restrict__ uncovered, uint32_t n_unc,
                                      const uint8_t* __restrict__ null_code) {
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) a.n_units * n_unc) return;
    const uint32_t unit_local = (uint32_t) (tid / n_unc), m = uncovered[tid % n_unc];
    const uint32_t code = null_code[m];
    if (code == 0) return;
    const uint32_t dst_row = a.dst_rows ? a.dst_rows[unit_local] : unit_local;
    uint32_t* dst = a.dst_base + (uint64_t) dst_row * a.row_words;
    for (uint32_t p = 0; p < a.planes; ++p) {
        if (!((code >> p) & 1u)) continue;
        uint32_t* out = dst + (uint64_t) (gam * a.planes + p) * a.plane_words + (m >> 5);
        atomicOr(out, 1u << (m & 31));
        if (a.doubled) atomicOr(out + a.planes * a.plane_words, 1u << (m & 31));
    }
}

// ------------------------------------------------------------------ native draws -> tape layout

// entry e of unit u: gamete = e >= n_chr0, chromosome = e - (gamete ? n_chr0 : 0)
__global__ void export_headers_kernel(const MeiosisArgs a, uint32_t* out_k, uint8_t* out_s, uint64_t* out_k64) {
    const uint32_t per_unit = a.map[0].n_chr + (a.gametes_per_unit > 1 ? a.map[1].n_chr : 0);
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) a.n_units * per_unit) return;
    const uint32_t unit_local = (uint32_t) (tid / per_unit), e = (uint32_t) (tid % per_unit);
    const uint32_t gam = e >= a.map[0].n_chr ? 1u : 0u;
    const uint32_t c = e - (gam ? a.map[0].n_chr : 0);
    const ChrInfo ci = a.map[gam].chr[c];
    uint32_t k, strand;
    draw_header(a.draws, ci, a.draws.first_unit + unit_local, gam, c, 0, k, strand);
    out_k[tid] = k;
    out_s[tid] = (uint8_t) strand;
    out_k64[tid] = k;
}

__global__ void export_positions_kernel(const MeiosisArgs a, const uint32_t* k_arr, const uint64_t* posoff, double* out_pos) {
    const uint32_t per_unit = a.map[0].n_chr + (a.gametes_per_unit > 1 ? a.map[1].n_chr : 0);
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) a.n_units * per_unit) return;
    const uint32_t unit_local = (uint32_t) (tid / per_unit), e = (uint32_t) (tid % per_unit);
    const uint32_t gam = e >= a.map[0].n_chr ? 1u : 0u;
    const uint32_t c = e - (gam ? a.map[0].n_chr : 0);
    const uint32_t k = k_arr[tid];
    double* dst = out_pos + posoff[tid];
    for (uint32_t j = 0; j < k; ++j) dst[j] = draw_position(a.draws, a.draws.first_unit + unit_local, gam, c, 0, j);
}

// ------------------------------------------------------------------ random parent pairs on the device
// gsc_helper_parentchooser_cross_randomly / _between with no usage cap (core.c:8364-8402,
// 8594-8631): parent a = round(u * (n1 - 1)); parent b likewise over its group, redrawn while
// it collides with a when both come from the same group (gsc_randomdraw_replacementrules,
// core.c:8556-8581).  Draws: the native stream with chromosome index 0xffffffff, indexed by cross.
__global__ void random_pairs_kernel(uint32_t n_crosses, uint32_t family_size, const uint32_t* __restrict__ members1, uint32_t n1,
                                    const uint32_t* __restrict__ members2, uint32_t n2, uint32_t distinct, DrawView draws,
                                    uint32_t first_out_row, uint32_t* __restrict__ p1_rows, uint32_t* __restrict__ p2_rows,
                                    uint32_t* __restrict__ picked1, uint32_t* __restrict__ picked2, uint32_t* __restrict__ out_rows) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_crosses) return;
    const uint64_t unit = draws.first_unit + i;
    uint4 r = draw_block(draws, unit, 0, 0xffffffffu, 0);
    const uint32_t a = (uint32_t) rint(u01(r.x, r.y) * (double) (n1 - 1));
    uint32_t b = (uint32_t) rint(u01(r.z, r.w) * (double) (n2 - 1));
    for (uint32_t blk = 1; distinct && b == a && blk < 4096; ++blk) {
        r = draw_block(draws, unit, 0, 0xffffffffu, blk);
        b = (uint32_t) rint(u01(r.x, r.y) * (double) (n2 - 1));
        if (b == a) b = (uint32_t) rint(u01(r.z, r.w) * (double) (n2 - 1));
    }
    const uint32_t ra = members1 ? members1[a] : a, rb = members2 ? members2[b] : b;    // null: the candidates are rows 0 .. n - 1
    for (uint32_t f = 0; f < family_size; ++f) {
        const uint64_t o = (uint64_t) i * family_size + f;
        p1_rows[o] = ra; p2_rows[o] = rb;
        if (picked1) { picked1[o] = a; picked2[o] = b; }
        if (out_rows) out_rows[o] = first_out_row + (uint32_t) o;
    }
}

// ------------------------------------------------------------------ clone

__global__ void copy_rows_kernel(uint32_t* store, const uint32_t* src_rows, const uint32_t* dst_rows, uint32_t n, uint64_t row_words) {
    const uint64_t per_row = row_words / 4;
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) n * per_row) return;
    const uint32_t i = (uint32_t) (tid / per_row);
    const uint64_t w = tid % per_row;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(store + (uint64_t) src_rows[i] * row_words) + w);
    reinterpret_cast<uint4*>(store + (uint64_t) dst_rows[i] * row_words)[w] = v;
}

// ------------------------------------------------------------------ host side

MapView view_of(const DeviceMap& m) {
    MapView v;
    v.chr = m.d_chr; v.dists = m.d_dists; v.bucket = m.d_bucket; v.marker_index = m.d_marker_index;
    v.chunk_chr = m.d_chunk_chr; v.chunk_start = m.d_chunk_start; v.n_chr = m.n_chr; v.n_chunks = m.n_chunks;
    return v;
}

uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

unsigned grid_for(uint64_t threads, unsigned block) { return (unsigned) ((threads + block - 1) / block); }

int get_map(gsb200_ctx* ctx, uint32_t index, const DeviceMap** out, const char* what) {
    GSB_REQUIRE(index < ctx->maps.size() && ctx->maps[index].present, GSB200_ERR_ARG, "%s: no recombination map at index %u", what, index);
    *out = &ctx->maps[index];
    return GSB200_OK;
}

int check_flag(gsb200_ctx* ctx, int* value) {
    GSB_CUDA_TRY(cudaMemcpyAsync(value, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (*value) GSB_CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    return GSB200_OK;
}

// A map with holes leaves the uncovered markers of a gamete at '\0' (core.c:75).  Make sure every
// such marker has a code for '\0' (this may widen the store, so it runs before anything reads the
// store's geometry) and keep the per-marker code array on the device.
int prepare_uncovered(gsb200_ctx* ctx, const DeviceMap& m) {
    if (m.covers_all) return GSB200_OK;
    for (uint32_t mk : m.h_uncovered) {
        uint32_t code;
        GSB_TRY(ensure_symbol(ctx, mk, 0, &code));
    }
    if (ctx->null_built_for != ctx->dict_version || !ctx->s_null.ptr) {
        std::vector<uint8_t> codes(ctx->n_markers, 0);
        for (uint32_t mk = 0; mk < ctx->n_markers; ++mk) {
            const uint8_t* sym = &ctx->dict_sym[(size_t) mk << ctx->planes];
            for (uint32_t c = 0; c < ctx->dict_n[mk]; ++c) if (sym[c] == 0) { codes[mk] = (uint8_t) c; break; }
        }
        GSB_TRY(ctx->s_null.ensure(ctx->n_markers));
        GSB_CUDA_TRY(cudaMemcpyAsync(ctx->s_null.ptr, codes.data(), ctx->n_markers, cudaMemcpyHostToDevice, ctx->stream));
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->null_built_for = ctx->dict_version;
    }
    return GSB200_OK;
}

__global__ void pool_wait_all_kernel(const PoolWait pw) {
    if (threadIdx.x < pw.world) {
        const long long t0 = clock64();
        while (pw.flags[threadIdx.x] < pw.epoch) {
            if (clock64() - t0 > kPoolWaitCycles) { atomicExch(pw.err, 1); break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

struct Batch {
    const char* what;
    uint32_t n_units;
    const uint32_t* src_base; const uint32_t* d_src0; const uint32_t* d_src1;
    uint32_t* dst_base; const uint32_t* d_dst;
    const DeviceMap* map0; const DeviceMap* map1;
    uint32_t gametes_per_unit, doubled;
    DrawView draws;
    uint32_t tape_xcap;        // TAPE: most crossovers in one gamete
    // fused GEBV (null: none); only taken on the fast kernel
    const long long* bv_tab = nullptr; const uint32_t* bv_slot = nullptr; const long long* bv_q = nullptr;
    unsigned long long* bv_acc = nullptr;
    uint32_t bv_shift = 5;
    const uint32_t* d_order = nullptr;   // gametes in processing order, or null
    const PoolWait* pw = nullptr;        // the source rows are a pool that is still arriving, or null
};

// Philox draws materialised in tape layout on the device (s_plan[4..7]).
int materialise_draws(gsb200_ctx* ctx, MeiosisArgs& a, uint64_t* n_positions_out) {
    const uint32_t per_unit = a.map[0].n_chr + (a.gametes_per_unit > 1 ? a.map[1].n_chr : 0);
    const uint64_t n_entries = (uint64_t) a.n_units * per_unit;
    GSB_TRY(ctx->s_plan[4].ensure(n_entries * sizeof(uint32_t)));
    GSB_TRY(ctx->s_plan[5].ensure(n_entries));
    GSB_TRY(ctx->s_plan[6].ensure((n_entries + 1) * sizeof(uint64_t)));
    uint32_t* d_k = (uint32_t*) ctx->s_plan[4].ptr;
    uint8_t* d_s = (uint8_t*) ctx->s_plan[5].ptr;
    uint64_t* d_off = (uint64_t*) ctx->s_plan[6].ptr;
    export_headers_kernel<<<grid_for(n_entries, 256), 256, 0, ctx->stream>>>(a, d_k, d_s, d_off);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    uint64_t last_k = 0, last_off = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&last_k, d_off + n_entries - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    size_t tmp_bytes = 0;
    GSB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_off, d_off, n_entries, ctx->stream));
    GSB_TRY(ctx->s_plan[3].ensure(tmp_bytes));
    GSB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(ctx->s_plan[3].ptr, tmp_bytes, d_off, d_off, n_entries, ctx->stream));
    ++ctx->launches;
    GSB_CUDA_TRY(cudaMemcpyAsync(&last_off, d_off + n_entries - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const uint64_t n_pos = last_off + last_k;
    GSB_TRY(ctx->s_plan[7].ensure(std::max<uint64_t>(n_pos, 1) * sizeof(double)));
    export_positions_kernel<<<grid_for(n_entries, 256), 256, 0, ctx->stream>>>(a, d_k, d_off, (double*) ctx->s_plan[7].ptr);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    a.draws.mode = GSB200_DRAWS_TAPE;
    a.draws.tape_k = d_k;
    a.draws.tape_s = d_s;
    a.draws.tape_posoff = d_off;
    a.draws.tape_pos = (const double*) ctx->s_plan[7].ptr;
    a.draws.unit_stride = per_unit;
    a.draws.base_offset = 0;
    *n_positions_out = n_pos;
    return GSB200_OK;
}

int run_general(gsb200_ctx* ctx, MeiosisArgs a, const DeviceMap* const* maps) {
    if (a.draws.mode == GSB200_DRAWS_PHILOX) {
        uint64_t n_pos = 0;
        GSB_TRY(materialise_draws(ctx, a, &n_pos));
    }
    zero_rows_kernel<<<grid_for((uint64_t) a.n_units * (a.row_words / 4), 256), 256, 0, ctx->stream>>>(
        a.dst_base, a.dst_rows, a.n_units, a.row_words);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    for (uint32_t gam = 0; gam < a.gametes_per_unit; ++gam) {
        const uint32_t n_unc = (uint32_t) maps[gam]->h_uncovered.size();
        if (n_unc) {
            fill_uncovered_kernel<<<grid_for((uint64_t) a.n_units * n_unc, 256), 256, 0, ctx->stream>>>(
                a, gam, maps[gam]->d_uncovered, n_unc, (const uint8_t*) ctx->s_null.ptr);
            ++ctx->launches;
            GSB_CUDA_TRY(cudaGetLastError());
        }
        const uint64_t threads = (uint64_t) a.n_units * a.map[gam].n_chunks;
        if (threads == 0) continue;
        splice_general_kernel<<<grid_for(threads, 128), 128, 0, ctx->stream>>>(a, gam);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
    }
    return GSB200_OK;
}

// Launch one batch of gametes. Asynchronous unless the fast kernel overflows its plan
// capacity (then the batch is redone on the general kernel) — `may_sync` says whether this
// call is allowed to wait for that verdict (the `_dev` entry points defer it).
int run_batch(gsb200_ctx* ctx, const Batch& b, bool may_sync) {
    if (b.n_units == 0) return GSB200_OK;
    MeiosisArgs a;
    a.src_base = b.src_base; a.src_rows0 = b.d_src0; a.src_rows1 = b.d_src1;
    a.dst_base = b.dst_base; a.dst_rows = b.d_dst;
    a.row_words = ctx->row_words(); a.plane_words = ctx->plane_words; a.planes = ctx->planes; a.n_markers = ctx->n_markers;
    a.n_units = b.n_units; a.gametes_per_unit = b.gametes_per_unit; a.doubled = b.doubled;
    a.fix_chunked = 0;
    a.map[0] = view_of(*b.map0);
    a.map[1] = view_of(b.gametes_per_unit > 1 ? *b.map1 : *b.map0);
    a.draws = b.draws;
    a.flag = ctx->d_flag;
    a.bv_tab = b.bv_tab; a.bv_slot = b.bv_slot; a.bv_q = b.bv_q; a.bv_acc = b.bv_acc; a.bv_shift = b.bv_shift;
    a.order = b.d_order;
    memset(&a.pw, 0, sizeof a.pw);
    if (b.pw) a.pw = *b.pw;

    const DeviceMap* maps[2] = { b.map0, b.gametes_per_unit > 1 ? b.map1 : b.map0 };
    if (b.draws.mode == GSB200_DRAWS_PHILOX) {
        for (const DeviceMap* m : maps)
            GSB_REQUIRE(m->lambda_max <= kMaxNativeLambda, GSB200_ERR_UNSUPPORTED,
                        "%s: a linkage group expects %.1f crossovers; the native draw stream supports at most %.0f",
                        b.what, m->lambda_max, kMaxNativeLambda);
        GSB_REQUIRE(b.draws.generation < kMaxNativeGenerations, GSB200_ERR_UNSUPPORTED,
                    "%s: at most %u chained generations per call", b.what, kMaxNativeGenerations);
    }

    // plan geometry of the fast kernel
    bool fast = maps[0]->flat && maps[1]->flat;
    a.n_u = ctx->plane_words / 4;
    a.parity_words = (a.n_u + 31) / 32;
    const uint32_t n_chr_max = std::max(maps[0]->n_chr, maps[1]->n_chr);
    a.chr_words = (n_chr_max + 31) / 32;
    uint32_t xcap;
    if (b.draws.mode == GSB200_DRAWS_TAPE) {
        xcap = b.tape_xcap;
    } else {
        const double ls = std::max(maps[0]->lambda_sum, maps[1]->lambda_sum);
        xcap = (uint32_t) std::ceil(ls + 8.0 * std::sqrt(ls) + 16.0);
    }
    xcap = std::max<uint32_t>((xcap + 31) / 32 * 32, 32);
    a.xcap = xcap;
    a.tcap = (xcap + 2 * n_chr_max + 1) / 2 * 2;
    a.warp_smem_words = (a.parity_words * (b.bv_acc ? 3 : 1) + 2 * a.chr_words + a.tcap + 2 * a.xcap + 3) / 4 * 4;
    if ((uint64_t) b.n_units * b.gametes_per_unit >= (1ull << 32)) fast = false;
    uint32_t warps = 8;
    const size_t smem_limit = 200 * 1024;
    while (warps > 1 && (size_t) warps * a.warp_smem_words * 4 > smem_limit / 2) warps >>= 1;
    if ((size_t) warps * a.warp_smem_words * 4 > smem_limit) fast = false;

    GSB_REQUIRE(fast || !b.bv_acc, GSB200_ERR_UNSUPPORTED, "%s: internal: fused GEBV needs the flat splice kernel", b.what);
    if (fast) {
        const size_t smem = (size_t) warps * a.warp_smem_words * 4;
        // resident blocks per SM the plain kernel is compiled for (registers per thread: 5 -> 48, 6 -> 40, 8 -> 32 with a few spills)
        static const int occ = [] { const char* v = getenv("GSB200_SPLICE_OCC"); return v ? atoi(v) : 5; }();
        auto* kernel = b.bv_acc ? splice_flat_kernel<true, 5> : occ == 8 ? splice_flat_kernel<false, 8> : occ == 6 ? splice_flat_kernel<false, 6> : splice_flat_kernel<false, 5>;
        if (smem > 48 * 1024)
            GSB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_limit));
        const uint64_t n_gametes = (uint64_t) a.n_units * a.gametes_per_unit;
        const uint64_t blocks = (n_gametes + warps - 1) / warps;
        {
            // what the resident warps (5 blocks per SM) read and write between a piece's copy and its fix-ups, against the L2
            static const int fix_mode = [] { const char* v = getenv("GSB200_SPLICE_FIX"); return v ? atoi(v) : -1; }();      // 0: never, k: every k groups
            const uint64_t resident = std::min<uint64_t>(n_gametes, (uint64_t) ctx->sm_count * 5 * warps);
            const uint64_t in_flight = resident * (uint64_t) a.plane_words * 4 * a.planes * (a.doubled ? 3 : 2);
            a.fix_chunked = fix_mode >= 0 ? (uint32_t) fix_mode : (in_flight > ((uint64_t) 96 << 20) ? 2u : 0u);      // groups of 2 KB per round of fix-ups
        }
        const unsigned grid = (unsigned) std::min<uint64_t>(blocks, (uint64_t) ctx->sm_count * 64);
        kernel<<<grid, warps * 32, smem, ctx->stream>>>(a);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
        if (!may_sync) return GSB200_OK;
        int flag = 0;
        GSB_TRY(check_flag(ctx, &flag));
        if (!flag) return GSB200_OK;
        // plan capacity exceeded somewhere in the batch: same draws, general kernel
    }
    GSB_REQUIRE(may_sync || !fast, GSB200_ERR_UNSUPPORTED, "%s: internal: unreachable", b.what);
    if (a.pw.flags) {                   // the general kernels do not wait per parent: everybody's rows first
        pool_wait_all_kernel<<<1, 32, 0, ctx->stream>>>(a.pw);
        ++ctx->launches;
    }
    return run_general(ctx, a, maps);
}

struct TapeUpload {
    const uint32_t* d_k = nullptr;
    const uint8_t* d_s = nullptr;
    const uint64_t* d_off = nullptr;
    const double* d_pos = nullptr;
    uint32_t xcap = 0;
};

// Validate a host tape against the request and put it on the device (s_plan[0..3]).
int upload_tape(gsb200_ctx* ctx, const gsb200_draws* d, uint64_t n_units, uint64_t entries_per_unit,
                const uint32_t* gamete_lengths, uint32_t gametes_in_unit, TapeUpload* out, const char* what) {
    const uint64_t n_entries = n_units * entries_per_unit;
    GSB_REQUIRE(d->n_entries == n_entries, GSB200_ERR_TAPE, "%s: the tape has %llu gamete-chromosome entries, the request needs %llu",
                what, (unsigned long long) d->n_entries, (unsigned long long) n_entries);
    GSB_REQUIRE(n_entries == 0 || (d->n_crossovers && d->start_strand), GSB200_ERR_TAPE, "%s: null tape arrays", what);
    std::vector<uint64_t> off(n_entries + 1);
    uint64_t run = 0;
    uint32_t xcap = 0;
    uint64_t e = 0;
    for (uint64_t u = 0; u < n_units; ++u) {
        for (uint32_t gI = 0; gI < gametes_in_unit; ++gI) {
            uint64_t in_gamete = 0;
            for (uint32_t c = 0; c < gamete_lengths[gI]; ++c, ++e) {
                off[e] = run;
                run += d->n_crossovers[e];
                in_gamete += d->n_crossovers[e];
            }
            GSB_REQUIRE(in_gamete < (1u << 30), GSB200_ERR_TAPE, "%s: absurd crossover count on the tape", what);
            xcap = std::max<uint32_t>(xcap, (uint32_t) in_gamete);
        }
    }
    off[n_entries] = run;
    GSB_REQUIRE(run == d->n_positions, GSB200_ERR_TAPE, "%s: the tape's counts sum to %llu positions but %llu were given",
                what, (unsigned long long) run, (unsigned long long) d->n_positions);
    GSB_REQUIRE(run == 0 || d->positions, GSB200_ERR_TAPE, "%s: null tape positions", what);
    GSB_TRY(ctx->s_plan[0].ensure(std::max<uint64_t>(n_entries, 1) * sizeof(uint32_t)));
    GSB_TRY(ctx->s_plan[1].ensure(std::max<uint64_t>(n_entries, 1)));
    GSB_TRY(ctx->s_plan[2].ensure((n_entries + 1) * sizeof(uint64_t)));
    GSB_TRY(ctx->s_plan[3].ensure(std::max<uint64_t>(run, 1) * sizeof(double)));
    if (n_entries) {
        GSB_CUDA_TRY(cudaMemcpyAsync(ctx->s_plan[0].ptr, d->n_crossovers, n_entries * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        GSB_CUDA_TRY(cudaMemcpyAsync(ctx->s_plan[1].ptr, d->start_strand, n_entries, cudaMemcpyHostToDevice, ctx->stream));
    }
    GSB_CUDA_TRY(cudaMemcpyAsync(ctx->s_plan[2].ptr, off.data(), (n_entries + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (run) GSB_CUDA_TRY(cudaMemcpyAsync(ctx->s_plan[3].ptr, d->positions, run * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // `off` is about to go out of scope
    out->d_k = (const uint32_t*) ctx->s_plan[0].ptr;
    out->d_s = (const uint8_t*) ctx->s_plan[1].ptr;
    out->d_off = (const uint64_t*) ctx->s_plan[2].ptr;
    out->d_pos = (const double*) ctx->s_plan[3].ptr;
    out->xcap = xcap;
    return GSB200_OK;
}

int make_draw_view(const gsb200_draws* d, DrawView* v, const char* what) {
    GSB_REQUIRE(d != nullptr, GSB200_ERR_ARG, "%s: null draws", what);
    GSB_REQUIRE(d->mode == GSB200_DRAWS_PHILOX || d->mode == GSB200_DRAWS_TAPE, GSB200_ERR_ARG, "%s: unknown draw mode %u", what, d->mode);
    memset(v, 0, sizeof *v);
    v->mode = d->mode;
    const uint64_t key = splitmix64(splitmix64(d->seed) ^ (d->stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    v->key0 = (uint32_t) key;
    v->key1 = (uint32_t) (key >> 32);
    v->first_unit = d->first_unit;
    return GSB200_OK;
}

struct FusedBv { const long long* tab; const uint32_t* slot; const long long* q; unsigned long long* acc; uint32_t shift; };

// key = the row a gamete is made from, value = the gamete (2 * offspring + which parent)
// (with a pool that is still arriving: the rank's own slots first, then its neighbours' in the order their rows arrive —
// rank - 1 pushes to this rank first, then rank - 2, ...; `slot_bits` bits of slot below the owner's turn)
__global__ void gamete_keys_kernel(uint32_t n_gametes, const uint32_t* __restrict__ p1, const uint32_t* __restrict__ p2, uint32_t p2_offset,
                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, const PoolWait pw, uint32_t slot_bits) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gametes) return;
    uint32_t k = (g & 1u) ? p2[g >> 1] + p2_offset : p1[g >> 1];
    if (pw.bounds) {
        uint32_t owner = 0;
        while (owner + 1 < pw.world && k >= pw.bounds[owner + 1]) ++owner;
        k |= ((pw.rank + pw.world - owner) % pw.world) << slot_bits;
    }
    keys[g] = k;
    vals[g] = g;
}

// A parent pool that does not stay in L2 beside the offspring stream: the gametes are processed sorted by parent, so that
// the ~2n / n_pool gametes of a parent are spliced by neighbouring warps at the same time and its row comes from DRAM once
// instead of once per gamete (configs[2]: 10 000 parents = 125 MB, 200 gametes each: read traffic 12.5 GB -> 0.13 GB per
// million offspring).  The draws are keyed by the global offspring number, so the order changes nothing in the result
// (tests/test_gpu_sharding.py).  Measured, 50 k SNPs (profiles/r2_sorted_gametes.txt): 100 MB pool 3.69 -> 3.00 ms per 800 k
// offspring (0.83 -> 1.02 of the HBM peak on the algorithmic bytes), 50 MB 1.64 -> 1.51 ms per 400 k, 25 MB 0.80 -> 0.78 ms
// per 200 k; the 12.5 MB pool of the headline step sits in L2 anyway and would only pay the sort's ~27 us.
// keys1 / keys2: per offspring, which of the n_keys candidate parents made gamete 0 / 1 (keys2 counted from key2_offset on).
// *d_order stays null when sorting is not worth it.
// st: the stream the sort runs on (null: the context stream).  On another stream the scratch must not move (memory that has just
// been allocated on the context stream is not that stream's yet): then *deferred is set and nothing is enqueued.
int gamete_order(gsb200_ctx* ctx, uint32_t n, const uint32_t* keys1, const uint32_t* keys2, uint32_t key2_offset, uint32_t n_keys,
                 const uint32_t** d_order, const PoolWait* pw = nullptr, cudaStream_t st = nullptr, bool* deferred = nullptr, bool free_ride = false) {
    *d_order = nullptr;
    if (deferred) *deferred = false;
    const bool aside = st != nullptr && st != ctx->stream;
    if (!st) st = ctx->stream;
    // Worth it when the DRAM reads it saves outweigh the sort (~30 us, mostly launches): the saving is the parents' half of
    // the algorithmic bytes times the share of them that misses L2 in numbered order — ~0 for a 12 MB pool, ~1 from 110 MB
    // on (measured: 0.2 at 25 MB, 0.4 at 50 MB, 0.9 at 100 MB).  GSB200_SORT_GAMETES_MIN_MB, read at every call so that
    // tests can switch it, replaces the estimate by a plain size threshold (0: always, huge: never).
    const size_t pool_bytes = (size_t) n_keys * ctx->row_words() * sizeof(uint32_t);
    if (n < 4 * (size_t) n_keys || n >= (1u << 30)) return GSB200_OK;
    if (const char* min_mb_env = getenv("GSB200_SORT_GAMETES_MIN_MB")) {
        if (pool_bytes < ((size_t) atol(min_mb_env) << 20)) return GSB200_OK;
    } else {
        const double pool_mb = (double) pool_bytes / (1 << 20);
        const double miss = std::min(1.0, std::max(0.0, (pool_mb - 12.0) / 100.0));
        double saved_us = (double) n * (double) (ctx->row_words() * sizeof(uint32_t) / 2) * miss / 6.5e6;      // at ~6.5 TB/s
        // a pool that is still arriving: sorted, the splice starts on the rank's own parents while the others' rows come in
        // over NVLink (~0.6 TB/s) instead of waiting for all of them
        if (pw && pw->world > 1) saved_us += (double) pool_bytes * (pw->world - 1) / pw->world / 0.6e6;
        // MEASURED: with the sort beside the GEBV kernel (free_ride) it pays on the device for ANY pool — neighbouring warps that
        // share their source rows help even a pool that sits in L2: 0.390 -> 0.372 ms per 100 k offspring from 1 000 parents,
        // step 0.739 -> 0.716 ms — but its enqueue (cub: key kernel + radix passes, ~25 us of host time) makes the host-bound
        // end-to-end path slower by more (0.770 -> 0.805 ms per generation; a one-launch cooperative counting sort: 0.83).
        // The break-even therefore stays where the in-line sort put it (GSB200_SORT_GAMETES_MIN_MB=0 sorts always).
        (void) free_ride;
        if (saved_us < 45.0 && !(pw && pw->world > 1)) return GSB200_OK;          // (shard.cu splits the push only when it is worth this)
    }
    const uint32_t n_g = 2 * n;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < n_keys) ++bits;
    const uint32_t slot_bits = (uint32_t) bits;
    PoolWait rot;
    memset(&rot, 0, sizeof rot);
    if (pw && pw->world > 1) {
        rot = *pw;
        while ((1u << (bits - (int) slot_bits)) < pw->world) ++bits;
        GSB_REQUIRE(bits <= 32, GSB200_ERR_ARG, "gamete order: too many parents for the sort key");
    }
    size_t temp_bytes = 0;
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint32_t*) nullptr, (uint32_t*) nullptr,
                                                 (const uint32_t*) nullptr, (uint32_t*) nullptr, (int) n_g, 0, bits, st));
    const size_t arr = ((size_t) n_g * sizeof(uint32_t) + 255) & ~(size_t) 255;
    if (aside && ctx->s_plan[0].bytes < 4 * arr + temp_bytes) {
        if (deferred) *deferred = true;
        GSB_TRY(ctx->s_plan[0].ensure(4 * arr + temp_bytes));      // grown now, in line: the next generation finds it in place
        return GSB200_OK;
    }
    GSB_TRY(ctx->s_plan[0].ensure(4 * arr + temp_bytes));
    uint32_t* k_in = (uint32_t*) ctx->s_plan[0].ptr;
    uint32_t* k_out = (uint32_t*) ((char*) k_in + arr);
    uint32_t* v_in = (uint32_t*) ((char*) k_in + 2 * arr);
    uint32_t* v_out = (uint32_t*) ((char*) k_in + 3 * arr);
    gamete_keys_kernel<<<grid_for(n_g, 256), 256, 0, st>>>(n_g, keys1, keys2, key2_offset, k_in, v_in, rot, slot_bits);
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs((char*) k_in + 4 * arr, temp_bytes, k_in, k_out, v_in, v_out, (int) n_g, 0, bits, st));
    ctx->launches += 3;
    GSB_CUDA_TRY(cudaGetLastError());
    *d_order = v_out;
    return GSB200_OK;
}

int cross_impl(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_p1, const uint32_t* d_p2, uint32_t map1, uint32_t map2,
               const uint32_t* d_out, const gsb200_draws* draws, bool may_sync, const char* what, const FusedBv* bv = nullptr,
               const uint32_t* src_base = nullptr /* rows the parents are read from; null: the store */,
               const uint32_t* d_order = nullptr /* gametes (2 * offspring + which) in processing order; null: as numbered */,
               const PoolWait* pw = nullptr) {
    const DeviceMap *m1, *m2;
    GSB_TRY(get_map(ctx, map1, &m1, what));
    GSB_TRY(get_map(ctx, map2, &m2, what));
    GSB_TRY(prepare_uncovered(ctx, *m1));
    GSB_TRY(prepare_uncovered(ctx, *m2));
    Batch b;
    b.what = what; b.n_units = n;
    b.src_base = src_base ? src_base : ctx->d_store; b.d_src0 = d_p1; b.d_src1 = d_p2; b.dst_base = ctx->d_store; b.d_dst = d_out;
    b.map0 = m1; b.map1 = m2; b.gametes_per_unit = 2; b.doubled = 0; b.tape_xcap = 0;
    if (bv) { b.bv_tab = bv->tab; b.bv_slot = bv->slot; b.bv_q = bv->q; b.bv_acc = bv->acc; b.bv_shift = bv->shift; }
    b.d_order = d_order;
    b.pw = pw;
    GSB_TRY(make_draw_view(draws, &b.draws, what));
    if (draws->mode == GSB200_DRAWS_TAPE) {
        GSB_REQUIRE(may_sync, GSB200_ERR_ARG, "%s: device-pointer entry points take PHILOX draws only", what);
        const uint32_t lens[2] = { m1->n_chr, m2->n_chr };
        TapeUpload t;
        GSB_TRY(upload_tape(ctx, draws, n, (uint64_t) m1->n_chr + m2->n_chr, lens, 2, &t, what));
        b.draws.tape_k = t.d_k; b.draws.tape_s = t.d_s; b.draws.tape_posoff = t.d_off; b.draws.tape_pos = t.d_pos;
        b.draws.unit_stride = (uint64_t) m1->n_chr + m2->n_chr;
        b.draws.base_offset = 0;
        b.draws.first_unit = 0;
        b.tape_xcap = t.xcap;
    }
    return run_batch(ctx, b, may_sync);
}

}  // namespace

namespace gsb {

int pool_random_cross(gsb200_ctx* ctx, const uint32_t* pool_base, uint32_t n_pool, uint32_t n, uint32_t map,
                      const uint32_t* d_out_rows, uint32_t first_out_row, const gsb200_draws* draws, uint32_t* d_picks,
                      const char* what, uint32_t* h_picks, cudaEvent_t picks_ready, const PoolWait* pw, cudaStream_t side_stream, cudaEvent_t side_may_start,
                      cudaEvent_t pairs_done) {
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(n_pool >= 2, GSB200_ERR_ARG, "%s: a pool of %u cannot be crossed at random", what, n_pool);
    GSB_REQUIRE(draws && draws->mode == GSB200_DRAWS_PHILOX, GSB200_ERR_ARG, "%s: needs PHILOX draws", what);
    GSB_REQUIRE(n < (1u << 31), GSB200_ERR_ARG, "%s: too many offspring in one call", what);
    GSB_REQUIRE(d_out_rows || (uint64_t) first_out_row + n <= ctx->capacity, GSB200_ERR_ARG, "%s: offspring rows beyond the store capacity", what);
    // s_rows[3]: parent 1 | parent 2 | offspring rows (when they are a contiguous run)
    const void* scratch_before = ctx->s_rows[3].ptr;
    GSB_TRY(ctx->s_rows[3].ensure((size_t) 4 * n * sizeof(uint32_t)));
    uint32_t* d_p1 = (uint32_t*) ctx->s_rows[3].ptr;
    uint32_t *d_p2 = d_p1 + n, *d_run = d_p2 + n;
    DrawView dv;
    GSB_TRY(make_draw_view(draws, &dv, what));
    // The pairs depend on the seed, the stream position and the SIZE of the pool only — not on who was selected — so with a side
    // stream they are drawn (and the picks sent down) as soon as the caller's event says the scratch is free: beside the
    // evaluation and selection queued in front of this call on the context stream, not behind them.  (Scratch that has just
    // been re-allocated on the context stream is not visible to another stream yet: then everything stays in line.)
    const bool early = side_stream && side_may_start && pairs_done && scratch_before == ctx->s_rows[3].ptr;
    cudaStream_t ps = early ? side_stream : ctx->stream;
    if (early) GSB_CUDA_TRY(cudaStreamWaitEvent(side_stream, side_may_start, 0));
    random_pairs_kernel<<<grid_for(n, 256), 256, 0, ps>>>(n, 1, nullptr, n_pool, nullptr, n_pool, 1u, dv, first_out_row, d_p1, d_p2,
                                                          d_picks, d_picks ? d_picks + n : nullptr, d_out_rows ? nullptr : d_run);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    // ... and so does the order the gametes are spliced in (sorted by parent SLOT when the pool outgrows L2): sorted over there too,
    // unless the order also depends on who owns which slot (a pool still arriving, pw)
    const uint32_t* d_order = nullptr;
    bool order_pending = true;
    if (early && !pw) {
        bool deferred = false;
        GSB_TRY(gamete_order(ctx, n, d_p1, d_p2, 0, n_pool, &d_order, nullptr, side_stream, &deferred, true));
        order_pending = deferred;
    }
    if (early) {
        GSB_CUDA_TRY(cudaEventRecord(pairs_done, side_stream));
        GSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, pairs_done, 0));
    }
    const bool picks_down = h_picks && d_picks;
    if (picks_down) {      // the picks leave before the splice is enqueued: the host has them while the offspring are being made
        GSB_CUDA_TRY(cudaMemcpyAsync(h_picks, d_picks, (size_t) 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ps));
        if (picks_ready) GSB_CUDA_TRY(cudaEventRecord(picks_ready, ps));
    }
    // A pool that no longer fits L2 beside the offspring stream (configs[2]: 10 000 parents = 125 MB) is re-read ~200 times
    // per generation; left to LRU it is flushed by the 12.5 GB the same kernel writes.  Part of L2 is therefore set aside
    // for the pool's lines for the duration of the splice (access-policy window on the stream: hits persist, everything
    // else streams).  MEASURED AND LEFT OFF: with the 79 MB set-aside the splice of 1 M offspring from 10 000 parents takes 12.4 ms
    // instead of 4.65 (misses streaming), 10.2 ms (misses normal), 4.66 ms with half the set-aside — the 12.5 GB write stream needs
    // the L2 more than the pool does.  GSB200_L2_PERSIST=1..4 switches the variants on again (profiles/scripts/config3_ab.py).
    static const int persist_mode = [] { const char* v = getenv("GSB200_L2_PERSIST"); return v ? atoi(v) : 0; }();
    const bool persist_on = persist_mode != 0;
    const size_t pool_bytes = (size_t) n_pool * ctx->row_words() * sizeof(uint32_t);
    bool window = false;
    if (persist_on && pool_bytes > ((size_t) 48 << 20) && ctx->l2_persist_max > 0 && pool_bytes <= ctx->l2_window_max) {
        const size_t want = persist_mode == 4 ? ctx->l2_persist_max / 2 : ctx->l2_persist_max;
        if (ctx->l2_persist_set != want) {
            GSB_CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            ctx->l2_persist_set = want;
            fprintf(stderr, "gsb200: persisting L2 set-aside %zu MB of max %zu MB, window max %zu MB\n", want >> 20, ctx->l2_persist_max >> 20, ctx->l2_window_max >> 20);
        }
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof attr);
        attr.accessPolicyWindow.base_ptr = const_cast<uint32_t*>(pool_base);
        attr.accessPolicyWindow.num_bytes = pool_bytes;
        attr.accessPolicyWindow.hitRatio = (float) std::min(1.0, (persist_mode == 3 ? 0.45 : 0.9) * (double) ctx->l2_persist_max / (double) pool_bytes);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = persist_mode == 1 ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
        GSB_CUDA_TRY(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        window = true;
    }
    if (order_pending) GSB_TRY(gamete_order(ctx, n, d_p1, d_p2, 0, n_pool, &d_order, pw));
    const int st = cross_impl(ctx, n, d_p1, d_p2, map, map, d_out_rows ? d_out_rows : d_run, draws, false, what, nullptr, pool_base, d_order, pw);
    if (window) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof attr);
        attr.accessPolicyWindow.num_bytes = 0;
        GSB_CUDA_TRY(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    }
    // joined behind the splice: whatever comes next on the context stream may reuse the picks' scratch
    if (early && picks_down && picks_ready) GSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, picks_ready, 0));
    return st;
}

}  // namespace gsb

extern "C" {

int gsb200_cross(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent1_rows, const uint32_t* parent2_rows,
                 uint32_t map1, uint32_t map2, const uint32_t* out_rows, const gsb200_draws* draws) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_TRY(check_rows(ctx, n, parent1_rows, "gsb200_cross (parent 1)"));
    GSB_TRY(check_rows(ctx, n, parent2_rows, "gsb200_cross (parent 2)"));
    GSB_TRY(check_rows(ctx, n, out_rows, "gsb200_cross (offspring)"));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t *d1, *d2, *dout;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], parent1_rows, n, &d1));
    GSB_TRY(upload_u32(ctx, ctx->s_rows[1], parent2_rows, n, &d2));
    GSB_TRY(upload_u32(ctx, ctx->s_rows[2], out_rows, n, &dout));
    GSB_TRY(cross_impl(ctx, n, d1, d2, map1, map2, dout, draws, true, "gsb200_cross"));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

int gsb200_cross_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_parent1_rows, const uint32_t* d_parent2_rows,
                     uint32_t map1, uint32_t map2, const uint32_t* d_out_rows, const gsb200_draws* draws) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(d_parent1_rows && d_parent2_rows && d_out_rows, GSB200_ERR_ARG, "gsb200_cross_dev: null row array");
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    return cross_impl(ctx, n, d_parent1_rows, d_parent2_rows, map1, map2, d_out_rows, draws, false, "gsb200_cross_dev");
}

int gsb200_cross_gebv_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_parent1_rows, const uint32_t* d_parent2_rows,
                          uint32_t map1, uint32_t map2, const uint32_t* d_out_rows, const gsb200_draws* draws,
                          uint32_t n_pool, const uint32_t* d_pool_rows, uint32_t eff_index, double* d_bv) {
    const char* what = "gsb200_cross_gebv_dev";
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(d_parent1_rows && d_parent2_rows && d_out_rows && d_bv, GSB200_ERR_ARG, "%s: null array", what);
    GSB_REQUIRE(draws != nullptr, GSB200_ERR_ARG, "%s: null draws", what);
    GSB_REQUIRE(eff_index < ctx->effects.size() && ctx->effects[eff_index].present, GSB200_ERR_ARG,
                "%s: no effect set at index %u", what, eff_index);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const DeviceMap *m1, *m2;
    GSB_TRY(get_map(ctx, map1, &m1, what));
    GSB_TRY(get_map(ctx, map2, &m2, what));
    // The plan-based GEBV needs the flat splice kernel, a one-plane store, the parents' pool and
    // tables that fit; anything else is the two separate passes.
    // one table entry per 8 markers while all tables together stay well inside L2, per 32 markers otherwise
    const uint32_t per_word = (size_t) n_pool * (4 * (size_t) ctx->plane_words + 2) * sizeof(long long) <= ((size_t) 64 << 20) ? 4u : 1u;
    const size_t tab_bytes = (size_t) n_pool * (per_word * (size_t) ctx->plane_words + 2) * sizeof(long long);
    // the fused kernel ranks the toggles of a gamete with 7-bit indexes: maps that expect anywhere near 128
    // toggles per gamete (crossovers + chromosome boundaries) take the two passes
    const double ls = std::max(m1->lambda_sum, m2->lambda_sum);
    const bool few_toggles = ls + 6.0 * std::sqrt(ls) + (double) std::max(m1->n_chr, m2->n_chr) <= 120.0;
    GSB_TRY(ensure_effect_tables(ctx, eff_index));
    const bool fused = ctx->planes == 1 && m1->flat && m2->flat && d_pool_rows && n_pool > 0 && tab_bytes <= ((size_t) 2 << 30)
                       && ctx->n_markers < (1u << 25) && few_toggles && draws->mode == GSB200_DRAWS_PHILOX
                       && ctx->effects[eff_index].all_finite;
    if (!fused) {
        GSB_TRY(cross_impl(ctx, n, d_parent1_rows, d_parent2_rows, map1, map2, d_out_rows, draws, false, what));
        return gsb200_gebv_dev(ctx, n, d_out_rows, eff_index, d_bv);
    }
    GSB_TRY(ensure_effect_tables(ctx, eff_index));
    const DeviceEffects& e = ctx->effects[eff_index];
    GSB_TRY(ctx->s_bv[0].ensure(tab_bytes));
    GSB_TRY(ctx->s_bv[1].ensure((size_t) ctx->capacity * sizeof(uint32_t)));
    GSB_TRY(ctx->s_bv[2].ensure((size_t) n * sizeof(long long)));
    long long* d_tab = (long long*) ctx->s_bv[0].ptr;
    uint32_t* d_slot = (uint32_t*) ctx->s_bv[1].ptr;
    unsigned long long* d_acc = (unsigned long long*) ctx->s_bv[2].ptr;
    // rows outside the pool keep slot 0xffffffff and raise the flag (gsb200_cross_dev_status); a stale slot
    // of an earlier call would silently give a wrong GEBV, so the map is reset every time (4 bytes per row)
    GSB_CUDA_TRY(cudaMemsetAsync(d_slot, 0xff, (size_t) ctx->capacity * sizeof(uint32_t), ctx->stream));
    GSB_CUDA_TRY(cudaMemsetAsync(d_acc, 0, (size_t) n * sizeof(long long), ctx->stream));
    // the per-row totals are accumulated with atomics: clear them first (the slot map reset covers the rest)
    GSB_CUDA_TRY(cudaMemsetAsync(d_tab, 0, tab_bytes, ctx->stream));
    {
        // enough blocks along the rows to fill the machine; every block reuses its Q slice for its rows
        const uint32_t word_blocks = (ctx->plane_words + kTabWords - 1) / kTabWords;
        uint32_t row_blocks = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t) ctx->sm_count * 16 / word_blocks + 1, (n_pool + kTabRows - 1) / kTabRows));
        row_blocks = std::min<uint32_t>(row_blocks, 65535);
        const uint32_t rows_per_block = ((n_pool + row_blocks - 1) / row_blocks + kTabRows - 1) / kTabRows * kTabRows;
        row_blocks = (n_pool + rows_per_block - 1) / rows_per_block;
        parent_words_kernel<<<dim3(word_blocks, row_blocks), kTabWords * kTabRows, 0, ctx->stream>>>(
            ctx->d_store, ctx->row_words(), ctx->plane_words, d_pool_rows, n_pool, rows_per_block, e.d_bvq, per_word, d_tab, d_slot);
        parent_scan_kernel<<<n_pool, 256, 0, ctx->stream>>>(per_word * ctx->plane_words, d_tab);
    }
    ctx->launches += 2;
    GSB_CUDA_TRY(cudaGetLastError());
    const FusedBv bv = { d_tab, d_slot, e.d_bvq, d_acc, per_word == 4 ? 3u : 5u };
    GSB_TRY(cross_impl(ctx, n, d_parent1_rows, d_parent2_rows, map1, map2, d_out_rows, draws, false, what, &bv));
    const double centre_sum = e.has_centre ? e.centre_sum : 0.0;
    bv_finish_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const long long*) d_acc, n, e.bvq_scale, e.base_all - centre_sum, d_bv);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB200_OK;
}

int gsb200_cross_random_pairs_begin(gsb200_ctx* ctx, uint32_t n_crosses, uint32_t family_size,
                              const uint32_t* member_rows1, uint32_t n1, const uint32_t* member_rows2, uint32_t n2,
                              uint32_t map1, uint32_t map2, const uint32_t* out_rows, uint32_t first_out_row,
                              const gsb200_draws* draws, uint32_t* picked1, uint32_t* picked2) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = "gsb200_cross_random_pairs";
    GSB_REQUIRE(!ctx->pending_cross.active, GSB200_ERR_ARG, "%s: the previous batch has not been ended", what);
    const uint64_t n64 = (uint64_t) n_crosses * family_size;
    if (n64 == 0) return GSB200_OK;
    GSB_REQUIRE(n64 < (1ull << 31), GSB200_ERR_ARG, "%s: too many offspring in one call", what);
    const uint32_t n = (uint32_t) n64;
    GSB_REQUIRE(draws && draws->mode == GSB200_DRAWS_PHILOX, GSB200_ERR_ARG, "%s: needs PHILOX draws", what);
    const bool same = member_rows2 == nullptr;
    GSB_REQUIRE(member_rows1 && n1 >= 1 && (same ? n1 >= 2 : n2 >= 1), GSB200_ERR_ARG, "%s: not enough candidate parents", what);
    GSB_TRY(check_rows(ctx, n1, member_rows1, what));
    if (!same) GSB_TRY(check_rows(ctx, n2, member_rows2, what));
    if (out_rows) GSB_TRY(check_rows(ctx, n, out_rows, what));
    else GSB_REQUIRE((uint64_t) first_out_row + n <= ctx->capacity, GSB200_ERR_ARG, "%s: offspring rows beyond the store capacity", what);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t *d_m1, *d_m2, *d_out = nullptr;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], member_rows1, n1, &d_m1));
    if (same) d_m2 = d_m1; else GSB_TRY(upload_u32(ctx, ctx->s_rows[1], member_rows2, n2, &d_m2));
    if (out_rows) GSB_TRY(upload_u32(ctx, ctx->s_rows[2], out_rows, n, &d_out));
    else { GSB_TRY(ctx->s_rows[2].ensure((size_t) n * sizeof(uint32_t))); d_out = (uint32_t*) ctx->s_rows[2].ptr; }
    GSB_TRY(ctx->s_rows[3].ensure((size_t) 4 * n * sizeof(uint32_t)));
    uint32_t* d_p1 = (uint32_t*) ctx->s_rows[3].ptr;
    uint32_t *d_p2 = d_p1 + n, *d_k1 = d_p2 + n, *d_k2 = d_k1 + n;
    DrawView dv;
    GSB_TRY(make_draw_view(draws, &dv, what));
    random_pairs_kernel<<<grid_for(n_crosses, 256), 256, 0, ctx->stream>>>(n_crosses, family_size, d_m1, n1, d_m2, same ? n1 : n2,
        same ? 1u : 0u, dv, first_out_row, d_p1, d_p2, d_k1, d_k2, out_rows ? nullptr : d_out);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    // the offspring of cross i are units i*family_size .. of the same stream
    gsb200_draws per_offspring = *draws;
    per_offspring.first_unit = draws->first_unit * family_size;
    // the picks leave before the splice kernel is enqueued: the host has them while the offspring are being made
    if (picked1) GSB_CUDA_TRY(cudaMemcpyAsync(picked1, d_k1, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (picked2) GSB_CUDA_TRY(cudaMemcpyAsync(picked2, d_k2, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (picked1 || picked2) GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // many offspring of a parent group that does not stay in L2: gametes sorted by parent (gamete_order above); d_k1 / d_k2
    // are the picks' positions in the candidate lists
    const uint32_t* d_order = nullptr;
    GSB_TRY(gamete_order(ctx, n, d_k1, d_k2, same ? 0u : n1, same ? n1 : n1 + n2, &d_order));
    GSB_TRY(cross_impl(ctx, n, d_p1, d_p2, map1, map2, d_out, &per_offspring, false, what, nullptr, nullptr, d_order));
    gsb200_ctx::PendingCross& pc = ctx->pending_cross;
    pc.active = true; pc.n = n; pc.map1 = map1; pc.map2 = map2;
    pc.d_p1 = d_p1; pc.d_p2 = d_p2; pc.d_out = d_out; pc.draws = per_offspring;
    return GSB200_OK;
}

int gsb200_cross_random_pairs_end(gsb200_ctx* ctx) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    gsb200_ctx::PendingCross& pc = ctx->pending_cross;
    if (!pc.active) return GSB200_OK;
    pc.active = false;
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    int flag = 0;
    GSB_TRY(check_flag(ctx, &flag));                 // synchronises
    if (flag)                                        // the on-chip plan overflowed somewhere: same parents, same draws, general kernel
        GSB_TRY(cross_impl(ctx, pc.n, pc.d_p1, pc.d_p2, pc.map1, pc.map2, pc.d_out, &pc.draws, true, "gsb200_cross_random_pairs"));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

int gsb200_cross_random_pairs(gsb200_ctx* ctx, uint32_t n_crosses, uint32_t family_size,
                              const uint32_t* member_rows1, uint32_t n1, const uint32_t* member_rows2, uint32_t n2,
                              uint32_t map1, uint32_t map2, const uint32_t* out_rows, uint32_t first_out_row,
                              const gsb200_draws* draws, uint32_t* picked1, uint32_t* picked2) {
    GSB_TRY(gsb200_cross_random_pairs_begin(ctx, n_crosses, family_size, member_rows1, n1, member_rows2, n2, map1, map2,
                                            out_rows, first_out_row, draws, picked1, picked2));
    return gsb200_cross_random_pairs_end(ctx);
}


int gsb200_cross_dev_status(gsb200_ctx* ctx) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    int flag = 0;
    GSB_TRY(check_flag(ctx, &flag));
    GSB_REQUIRE(flag != 4, GSB200_ERR_ARG, "gsb200_cross_gebv_dev: a parent row was not in the pool; the GEBVs of that batch are void");
    GSB_REQUIRE(flag == 0, GSB200_ERR_UNSUPPORTED,
                "a gsb200_cross_dev batch exceeded the splice plan's capacity; redo it with gsb200_cross");
    return GSB200_OK;
}

int gsb200_doubled_haploid(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent_rows, uint32_t map,
                           const uint32_t* out_rows, const gsb200_draws* draws) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    const char* what = "gsb200_doubled_haploid";
    GSB_TRY(check_rows(ctx, n, parent_rows, what));
    GSB_TRY(check_rows(ctx, n, out_rows, what));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const DeviceMap* m;
    GSB_TRY(get_map(ctx, map, &m, what));
    GSB_TRY(prepare_uncovered(ctx, *m));
    uint32_t *d1, *dout;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], parent_rows, n, &d1));
    GSB_TRY(upload_u32(ctx, ctx->s_rows[2], out_rows, n, &dout));
    Batch b;
    b.what = what; b.n_units = n;
    b.src_base = ctx->d_store; b.d_src0 = d1; b.d_src1 = d1; b.dst_base = ctx->d_store; b.d_dst = dout;
    b.map0 = m; b.map1 = m; b.gametes_per_unit = 1; b.doubled = 1; b.tape_xcap = 0;
    GSB_TRY(make_draw_view(draws, &b.draws, what));
    if (draws->mode == GSB200_DRAWS_TAPE) {
        const uint32_t lens[1] = { m->n_chr };
        TapeUpload t;
        GSB_TRY(upload_tape(ctx, draws, n, m->n_chr, lens, 1, &t, what));
        b.draws.tape_k = t.d_k; b.draws.tape_s = t.d_s; b.draws.tape_posoff = t.d_off; b.draws.tape_pos = t.d_pos;
        b.draws.unit_stride = m->n_chr;
        b.draws.first_unit = 0;
        b.tape_xcap = t.xcap;
    }
    GSB_TRY(run_batch(ctx, b, true));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

int gsb200_self_n_times(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent_rows, uint32_t n_generations,
                        uint32_t map, const uint32_t* out_rows, const gsb200_draws* draws) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = "gsb200_self_n_times";
    GSB_REQUIRE(n_generations >= 1, GSB200_ERR_ARG, "%s: n_generations must be at least 1", what);
    if (n == 0) return GSB200_OK;
    GSB_TRY(check_rows(ctx, n, parent_rows, what));
    GSB_TRY(check_rows(ctx, n, out_rows, what));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const DeviceMap* m;
    GSB_TRY(get_map(ctx, map, &m, what));
    GSB_TRY(prepare_uncovered(ctx, *m));
    uint32_t *d1, *dout;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], parent_rows, n, &d1));
    GSB_TRY(upload_u32(ctx, ctx->s_rows[2], out_rows, n, &dout));
    DrawView dv;
    GSB_TRY(make_draw_view(draws, &dv, what));
    TapeUpload t;
    const uint64_t per_unit = (uint64_t) n_generations * 2 * m->n_chr;
    if (draws->mode == GSB200_DRAWS_TAPE) {
        std::vector<uint32_t> lens(2 * (size_t) n_generations, m->n_chr);
        GSB_TRY(upload_tape(ctx, draws, n, per_unit, lens.data(), (uint32_t) lens.size(), &t, what));
    }
    // Intermediate generations live in two ping-pong scratch row sets, never in the store
    // (the reference ping-pongs a temporary genotype the same way, core.c:8908-8923); lines
    // are processed in slabs so that the scratch stays bounded.
    const uint64_t row_bytes = ctx->row_words() * sizeof(uint32_t);
    const uint32_t slab = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(n, ((uint64_t) 4 << 30) / row_bytes));
    if (n_generations > 1) {
        GSB_TRY(ctx->s_chain[0].ensure((uint64_t) slab * row_bytes));
        if (n_generations > 2) GSB_TRY(ctx->s_chain[1].ensure((uint64_t) slab * row_bytes));
    }
    for (uint32_t done = 0; done < n; done += slab) {
        const uint32_t cnt = std::min(slab, n - done);
        for (uint32_t gen = 0; gen < n_generations; ++gen) {
            Batch b;
            b.what = what; b.n_units = cnt;
            const bool first = gen == 0, last = gen + 1 == n_generations;
            // generation `gen` reads scratch (gen-1)&1 and writes scratch gen&1; the first reads the
            // store, the last writes it
            b.src_base = first ? ctx->d_store : (const uint32_t*) ctx->s_chain[(gen - 1) & 1].ptr;
            b.d_src0 = b.d_src1 = first ? d1 + done : nullptr;
            b.dst_base = last ? ctx->d_store : (uint32_t*) ctx->s_chain[gen & 1].ptr;
            b.d_dst = last ? dout + done : nullptr;
            b.map0 = b.map1 = m; b.gametes_per_unit = 2; b.doubled = 0; b.tape_xcap = t.xcap;
            b.draws = dv;
            b.draws.generation = gen;
            b.draws.first_unit = dv.first_unit + done;
            if (draws->mode == GSB200_DRAWS_TAPE) {
                b.draws.tape_k = t.d_k + (uint64_t) done * per_unit;
                b.draws.tape_s = t.d_s + (uint64_t) done * per_unit;
                b.draws.tape_posoff = t.d_off + (uint64_t) done * per_unit;
                b.draws.tape_pos = t.d_pos;
                b.draws.unit_stride = per_unit;
                b.draws.base_offset = (uint64_t) gen * 2 * m->n_chr;
                b.draws.first_unit = 0;
            }
            GSB_TRY(run_batch(ctx, b, true));
        }
    }
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

int gsb200_clone(gsb200_ctx* ctx, uint32_t n, const uint32_t* parent_rows, const uint32_t* out_rows) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_TRY(check_rows(ctx, n, parent_rows, "gsb200_clone"));
    GSB_TRY(check_rows(ctx, n, out_rows, "gsb200_clone"));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t *d1, *dout;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], parent_rows, n, &d1));
    GSB_TRY(upload_u32(ctx, ctx->s_rows[2], out_rows, n, &dout));
    copy_rows_kernel<<<grid_for((uint64_t) n * (ctx->row_words() / 4), 256), 256, 0, ctx->stream>>>(
        ctx->d_store, d1, dout, n, ctx->row_words());
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

int gsb200_export_draws(gsb200_ctx* ctx, uint32_t map, uint64_t n_units, uint32_t gametes_per_unit,
                        const gsb200_draws* philox, uint32_t* n_crossovers, uint8_t* start_strand,
                        double* positions, uint64_t positions_cap, uint64_t* n_positions) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    const char* what = "gsb200_export_draws";
    GSB_REQUIRE(philox && philox->mode == GSB200_DRAWS_PHILOX, GSB200_ERR_ARG, "%s: needs PHILOX draws", what);
    GSB_REQUIRE(gametes_per_unit == 1 || gametes_per_unit == 2, GSB200_ERR_ARG, "%s: gametes_per_unit must be 1 or 2", what);
    GSB_REQUIRE(n_units > 0 && n_units < (1ull << 31) && n_crossovers && start_strand && n_positions, GSB200_ERR_ARG, "%s: bad arguments", what);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const DeviceMap* m;
    GSB_TRY(get_map(ctx, map, &m, what));
    GSB_REQUIRE(m->lambda_max <= kMaxNativeLambda, GSB200_ERR_UNSUPPORTED, "%s: lambda too large for the native stream", what);
    MeiosisArgs a;
    memset(&a, 0, sizeof a);
    a.n_units = (uint32_t) n_units;
    a.gametes_per_unit = gametes_per_unit;
    a.map[0] = a.map[1] = view_of(*m);
    GSB_TRY(make_draw_view(philox, &a.draws, what));
    uint64_t n_pos = 0;
    GSB_TRY(materialise_draws(ctx, a, &n_pos));
    const uint64_t n_entries = n_units * gametes_per_unit * m->n_chr;
    *n_positions = n_pos;
    GSB_CUDA_TRY(cudaMemcpyAsync(n_crossovers, a.draws.tape_k, n_entries * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaMemcpyAsync(start_strand, a.draws.tape_s, n_entries, cudaMemcpyDeviceToHost, ctx->stream));
    if (positions && n_pos <= positions_cap && n_pos)
        GSB_CUDA_TRY(cudaMemcpyAsync(positions, a.draws.tape_pos, n_pos * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    GSB_REQUIRE(n_pos <= positions_cap || positions == nullptr, GSB200_ERR_ARG,
                "%s: %llu positions do not fit the buffer of %llu", what, (unsigned long long) n_pos, (unsigned long long) positions_cap);
    return GSB200_OK;
}

}  // extern "C"
