// local_tc5.cu — local (marker-block) GEBVs for MANY effect sets on the 5th-generation tensor cores:
// the dense-contraction path of BASELINE.json configs[4] (gsc_calculate_utility_local_bvs,
// core.c:9979-10090, computed for up to 10 effect sets in one pass over the genotypes).
//
// The contraction: (strands x markers) bits  x  (markers x 8 * sets) fixed-point digits, int32
// accumulation per marker block, recombined in fp64 per (strand, block, set).  gebv_tc5.cu showed
// that tcgen05.mma with N = 8 is bound by streaming its A operand (~90 cycles per 128 x 32-byte tile
// whatever N is); with N = 8 * sets the same operand is reused by every effect set — ten sets cost
// what one does.
//
//   * a CTA owns a marker SEGMENT — up to ~200 KB of B digits resident in shared memory — and walks
//     tiles of 64 individuals (128 strand rows = the 128 TMEM lanes);
//   * 4 expander warps: thread t loads 32 contiguous bytes (8 marker words) of ITS strand row, turns
//     each word into 8 operand registers `w & (0x01010101 << s)` (byte b = bit 8b + s at weight 2^s; the
//     weight is divided out of the digits on the host) and writes them to its own TMEM lane with
//     tcgen05.st.32x32b — one K = 32 step per word, so a step is 32 CONTIGUOUS markers and block
//     boundaries cost at most one extra step (the word that straddles them, its B digits zeroed outside
//     the block), exactly the step schedule of the mma.sync kernel (evaluate.cu);
//   * 1 issuer thread: per step one tcgen05.mma.kind::i8, M = 128, N = 8 * sets, K = 32, A from TMEM,
//     B through a shared-memory descriptor; the accumulator switches (double-buffered) at every block
//     boundary;
//   * 4 epilogue warps: tcgen05.ld of a finished block's 8 * sets digit sums per strand, Horner in fp64,
//     scale, add to the output matrices (fp64 atomics: a block may span two segments; the outputs are
//     zeroed by the caller and the per-block base is added by the segment the block starts in).
//
// Every mbarrier wait is bounded (tc5_common.cuh): a protocol bug raises the context's flag instead of
// hanging the GPU.
#include "gsb200_internal.cuh"
#include "tc5_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace gsb;
using namespace gsb::tc5;

namespace gsb {

namespace {

constexpr uint32_t kExpanders = 128, kEpilogue = 128;
constexpr uint32_t kThreads = kExpanders + kEpilogue + 32;
constexpr uint32_t kAStages = 4;
constexpr uint32_t kStageCols = 64;            // 8 words x 8 columns
constexpr uint32_t kDCol1 = 128;               // second accumulator
constexpr uint32_t kACol0 = 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kMaxSegSteps = 256;
constexpr uint32_t kBBytes = 192 * 1024;         // B digits of a segment (+ 256 bytes of slack)
constexpr uint32_t kTraceLen = 96;               // stages / flushes of CTA 0 whose timestamps GSB200_LOCAL_TRACE records
constexpr uint32_t kMaxRunBases = 1536;          // doubles of per-run bases kept in shared memory (runs x sets)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// four marker words of this thread's strand row -> 4 K-steps of operand bytes in its TMEM lane
__device__ __forceinline__ void store_words(uint32_t taddr, const uint4 w) {
    uint32_t r[32];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const uint32_t m = 0x01010101u << s;
        r[s] = w.x & m; r[8 + s] = w.y & m; r[16 + s] = w.z & m; r[24 + s] = w.w & m;
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

struct LocalTc5Params {
    const uint4* b[kLocalTc5MaxSets];      // per set: [n_steps][16] uint4 = 256 bytes per step, UMMA K-major layout
    const double* base[kLocalTc5MaxSets];  // [n_blocks]
    double* out[kLocalTc5MaxSets];         // [(2n) x n_blocks], zeroed
    double scale[kLocalTc5MaxSets];
};

// steps[s] = (word, block | first_step_of_block << 31)
__global__ void __launch_bounds__(kThreads, 1) local_tc5_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ rows, uint32_t n, uint32_t n_blocks,
        const uint2* __restrict__ steps, uint32_t n_steps, const uint32_t* __restrict__ seg_off /* [n_segments + 1] first step of every segment */,
        uint32_t n_tiles, uint32_t n_items, uint32_t items_per_cta, uint32_t n_sets, const LocalTc5Params P, int* __restrict__ flag,
        uint32_t dbg /* GSB200_LOCAL_DEBUG bits, measurements only: 1 no operand stores, 2 no mma, 4 no epilogue work, 8 no row loads */,
        long long* __restrict__ trace /* debugging: clock64 stamps of CTA 0's first kTraceLen stages per role, or null */) {
    extern __shared__ __align__(128) uint4 s_b[];                    // seg_steps x n_sets x 256 bytes
    __shared__ __align__(8) uint64_t s_full[kAStages], s_free[kAStages], s_dready[2], s_dfree[2];
    __shared__ uint2 s_steps[kMaxSegSteps];
    // the segment's runs of steps that belong to one block = its accumulators, in order: (block | starts-here << 31, first step)
    __shared__ uint2 s_runs[kMaxSegSteps];
    __shared__ double s_base[kMaxRunBases];          // [run][set]: the block's base where the block starts in this segment, else 0
    // per step, for the issuer: word & 7 | 8 if the step opens an operand stage (octet) | 16 if it opens an accumulator
    // (block) | octets to advance << 8 — precomputed so that the single issuing thread does next to nothing per MMA
    __shared__ uint32_t s_rec[kMaxSegSteps + 1];
    __shared__ uint32_t s_n_runs;
    __shared__ uint32_t s_tmem;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t i = 0; i < kAStages; ++i) { mbar_init(&s_full[i], kExpanders / 32); mbar_init(&s_free[i], 1); }
        for (uint32_t i = 0; i < 2; ++i) { mbar_init(&s_dready[i], 1); mbar_init(&s_dfree[i], kEpilogue / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t bar_full = smem_u32(s_full), bar_free = smem_u32(s_free);
    const uint32_t bar_dready = smem_u32(s_dready), bar_dfree = smem_u32(s_dfree);
    const uint32_t b_base = smem_u32(s_b);
    // M = 128 wants N % 16 == 0: an odd number of sets contracts 8 more columns against the bytes that follow
    // (the next step's first set, or the slack behind the last step) into accumulator columns nobody reads
    const uint32_t n_cols = (8 * n_sets + 15) & ~15u, step_bytes = 256 * n_sets;
    const uint32_t idesc = idesc_i8(n_cols);
    bool ok = true;

    const uint32_t item_lo = min(blockIdx.x * items_per_cta, n_items), item_hi = min(item_lo + items_per_cta, n_items);
    // running counts over the CTA's whole life: every role walks the same sequence, so each derives the
    // barrier phases by itself
    uint32_t q_oct = 0;          // operand stages (octets of 8 words) produced / consumed so far
    uint32_t q_flush = 0;        // accumulators finished so far

    for (uint32_t item = item_lo; item < item_hi;) {
        const uint32_t seg = item / n_tiles, run_end = min(item_hi, (seg + 1) * n_tiles);
        const uint32_t s_lo = seg_off[seg], s_hi = seg_off[seg + 1], n_seg = s_hi - s_lo;
        const uint32_t tile_lo = item - seg * n_tiles, n_run = run_end - item;

        // ---- this segment's step list and B digits -> shared memory (the previous segment's mma have all
        //      completed: the epilogue warps waited for its last accumulator before coming here)
        tc_fence_before();
        __syncthreads();
        for (uint32_t i = tid; i < n_seg; i += kThreads) s_steps[i] = steps[s_lo + i];
        for (uint32_t i = tid; i < n_seg * n_sets * 16; i += kThreads) {
            const uint32_t st = i / (n_sets * 16), r = i % (n_sets * 16), q = r >> 4, k = r & 15;
            s_b[i] = __ldg(P.b[q] + (size_t) (s_lo + st) * 16 + k);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to the tensor core's proxy
        __syncthreads();
        for (uint32_t st = tid; st <= n_seg; st += kThreads) {
            uint32_t rec = 24u;                                  // the sentinel behind the last step ends the issuer's fast loop
            if (st < n_seg) {
                const uint32_t oct = s_steps[st].x >> 3, blk = s_steps[st].y & 0x7fffffffu;
                const uint32_t adv = st == 0 ? 1u : oct - (s_steps[st - 1].x >> 3);
                rec = (s_steps[st].x & 7u) | (adv ? 8u : 0u) | ((st == 0 || blk != (s_steps[st - 1].y & 0x7fffffffu)) ? 16u : 0u) | (adv << 8);
            }
            s_rec[st] = rec;
        }
        if (warp == 8) {
            // the runs, by one warp: a step opens a run when its block differs from the previous step's
            uint32_t n_runs = 0;
            for (uint32_t s0 = 0; s0 < n_seg; s0 += 32) {
                const uint32_t st = s0 + lane;
                const bool opens = st < n_seg && (st == 0 || (s_steps[st].y & 0x7fffffffu) != (s_steps[st - 1].y & 0x7fffffffu));
                const uint32_t ballot = __ballot_sync(0xffffffffu, opens);
                if (opens) s_runs[n_runs + __popc(ballot & ((1u << lane) - 1u))] = make_uint2(s_steps[st].y, st);
                n_runs += __popc(ballot);
            }
            if (lane == 0) s_n_runs = n_runs;
        }
        __syncthreads();
        const uint32_t n_runs = s_n_runs;
        // does the segment's first run start its block, does its last run end it?  (every other run is a whole block)
        const bool last_run_ends = s_hi == n_steps || (steps[s_hi].y & 0x7fffffffu) != (s_steps[n_seg - 1].y & 0x7fffffffu);
        const bool bases_in_smem = n_runs * n_sets <= kMaxRunBases;
        if (bases_in_smem)
            for (uint32_t i = tid; i < n_runs * n_sets; i += kThreads) {
                const uint2 run = s_runs[i / n_sets];
                s_base[i] = (run.x >> 31) ? P.base[i % n_sets][run.x & 0x7fffffffu] : 0.0;
            }
        __syncthreads();
        tc_fence_after();
        const uint32_t o_first = s_steps[0].x >> 3, o_last = s_steps[n_seg - 1].x >> 3, n_oct = o_last - o_first + 1;

        if (warp < 4) {
            // ---------------- expanders: thread tid = strand row tid of the tile
            auto row_ptr = [&](uint32_t tile) -> const uint4* {
                const uint32_t ind = min(min(tile, n_tiles - 1) * 64 + (tid >> 1), n - 1);
                return reinterpret_cast<const uint4*>(store + (uint64_t) rows[ind] * row_words + (tid & 1) * plane_words) + 2 * o_first;
            };
            const uint32_t total = n_run * n_oct;
            constexpr uint32_t kDepth = 4;
            uint4 buf[kDepth][2];
            uint32_t l_tile = tile_lo, l_oct = 0;
            const uint4* lptr = row_ptr(l_tile);
            auto request = [&](uint4 (&b)[2]) {
                if (!(dbg & 8u)) { b[0] = ldg_stream(lptr + 2 * l_oct); b[1] = ldg_stream(lptr + 2 * l_oct + 1); }
                else b[0] = b[1] = make_uint4(l_oct, tid, 3u, 4u);
                if (++l_oct == n_oct) { l_oct = 0; ++l_tile; lptr = row_ptr(l_tile); }
            };
#pragma unroll
            for (uint32_t k = 0; k < kDepth; ++k) if (k < total) request(buf[k]);
            auto produce = [&](const uint4 (&b)[2]) {
                const uint32_t stage = q_oct % kAStages, use = q_oct / kAStages;
                const bool tr = trace && blockIdx.x == 0 && tid == 0 && q_oct < kTraceLen;
                if (tr) trace[0 * kTraceLen + q_oct] = clock64();
                if (use > 0) {                                 // the mma that read this stage kAStages octets ago
                    mbar_wait(bar_free + 8 * stage, (use - 1) & 1, ok);
                    tc_fence_after();
                }
                const uint32_t taddr = tmem + ((warp * 32) << 16) + kACol0 + kStageCols * stage;
                if (tr) trace[1 * kTraceLen + q_oct] = clock64();
                if (!(dbg & 1u)) store_words(taddr, b[0]);
                if (tr) trace[2 * kTraceLen + q_oct] = clock64();
                if (!(dbg & 1u)) store_words(taddr + 32, b[1]);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                if (tr) trace[3 * kTraceLen + q_oct] = clock64();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * stage);
                ++q_oct;
            };
            for (uint32_t c0 = 0; c0 < total; c0 += kDepth) {
#pragma unroll
                for (uint32_t k = 0; k < kDepth; ++k) {
                    if (c0 + k < total) {
                        produce(buf[k]);
                        if (c0 + k + kDepth < total) request(buf[k]);
                    }
                }
            }
        } else if (warp < 8) {
            // ---------------- epilogue: thread (tid - 128) = TMEM lane = strand row of the tile
            const uint32_t t = tid - kExpanders, wq = warp - 4;
            for (uint32_t it = 0; it < n_run; ++it) {
                const uint32_t tile = tile_lo + it;
                const uint64_t out_row = (uint64_t) tile * 128 + t;          // = 2 * individual + strand
                const bool live = (tile * 64 + (t >> 1)) < n;
                for (uint32_t run = 0; run < n_runs; ++run) {
                    const uint32_t blk = s_runs[run].x & 0x7fffffffu, starts = s_runs[run].x >> 31;
                    const bool whole = starts && (run + 1 < n_runs || last_run_ends);      // complete sums: plain stores, no atomics
                    const uint32_t dbuf = q_flush & 1;
                    const bool tr = trace && blockIdx.x == 0 && t == 0 && q_flush < kTraceLen;
                    if (tr) trace[6 * kTraceLen + q_flush] = clock64();
                    mbar_wait(bar_dready + 8 * dbuf, (q_flush >> 1) & 1, ok);
                    tc_fence_after();
                    if (tr) trace[7 * kTraceLen + q_flush] = clock64();
                    const uint32_t taddr = tmem + ((wq * 32) << 16) + (dbuf ? kDCol1 : 0u);
                    double* out = P.out[0] + out_row * n_blocks + blk;
                    for (uint32_t q0 = 0; q0 < ((dbg & 4u) ? 0u : n_sets); q0 += 4) {
                        // four sets at a time: the loads first, then four independent recombinations
                        uint32_t d[4][8];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(d[j][0]), "=r"(d[j][1]), "=r"(d[j][2]), "=r"(d[j][3]), "=r"(d[j][4]), "=r"(d[j][5]), "=r"(d[j][6]), "=r"(d[j][7])
                                         : "r"(taddr + 8 * min(q0 + j, n_sets - 1)) : "memory");
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t q = q0 + j;
                            // sum_k d_k 256^k: the two halves are exact in int64 (|d_k| < 2^31), one fp64 fma joins them
                            const long long lo = (long long) (int) d[j][0] + ((long long) (int) d[j][1] << 8) + ((long long) (int) d[j][2] << 16) + ((long long) (int) d[j][3] << 24);
                            const long long hi = (long long) (int) d[j][4] + ((long long) (int) d[j][5] << 8) + ((long long) (int) d[j][6] << 16) + ((long long) (int) d[j][7] << 24);
                            if (q < n_sets) {
                                double v = fma((double) hi, 4294967296.0, (double) lo) * P.scale[q];
                                if (starts) v += bases_in_smem ? s_base[run * n_sets + q] : P.base[q][blk];
                                if (live) { if (whole) P.out[q][out - P.out[0]] = v; else atomicAdd(P.out[q] + (out - P.out[0]), v); }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_dfree + 8 * dbuf);
                    if (tr) trace[8 * kTraceLen + q_flush] = clock64();
                    ++q_flush;
                }
            }
        } else {
            // ---------------- issuer: one thread.  A lone thread retires an instruction every ~4 cycles, so the
            // step loop is kept to a dozen instructions (record prefetch, operand column, descriptor bump, mma) —
            // a first version with the stage / accumulator bookkeeping inline spent ~300 cycles per step, three
            // times what the mma itself takes (profiles/r2_local_tc5.txt).  Octet and block changes, flagged in
            // the precomputed records, leave the fast loop.
            if (lane == 0) {
                const uint64_t desc0 = b_descriptor(b_base);
                const uint32_t desc_step = step_bytes >> 4;          // the start-address field counts 16-byte units; all of shared memory fits its 14 bits
                for (uint32_t it = 0; it < n_run; ++it) {
                    uint32_t stage = 0, dbuf = 0, a_base = 0, d_addr = 0, s = 0;
                    bool have_oct = false, have_blk = false;
                    uint64_t desc = desc0;
                    uint32_t rec = s_rec[0];
                    while (s < n_seg) {
                        uint32_t acc = 1;
                        if (rec & 8u) {
                            // (a gap between two blocks may skip whole octets; the expanders deliver every octet of the segment)
                            for (uint32_t adv = rec >> 8; adv; --adv) {
                                if (have_oct) tc_commit(bar_free + 8 * stage);
                                have_oct = true;
                                stage = q_oct % kAStages;
                                if (trace && blockIdx.x == 0 && q_oct < kTraceLen) trace[4 * kTraceLen + q_oct] = clock64();
                                mbar_wait(bar_full + 8 * stage, (q_oct / kAStages) & 1, ok);
                                tc_fence_after();
                                if (trace && blockIdx.x == 0 && q_oct < kTraceLen) trace[5 * kTraceLen + q_oct] = clock64();
                                ++q_oct;
                            }
                            a_base = tmem + kACol0 + kStageCols * stage;
                        }
                        if (rec & 16u) {
                            if (have_blk) { tc_commit(bar_dready + 8 * dbuf); ++q_flush; }
                            have_blk = true;
                            dbuf = q_flush & 1;
                            if (q_flush >= 2) {                      // the epilogue has read this accumulator's previous sums
                                mbar_wait(bar_dfree + 8 * dbuf, ((q_flush >> 1) - 1) & 1, ok);
                                tc_fence_after();
                            }
                            d_addr = tmem + (dbuf ? kDCol1 : 0u);
                            acc = 0;
                        }
                        // the steps up to the next octet or block change (the record behind the last step has both flags set)
                        do {
                            if (!(dbg & 2u)) umma_i8_ts(d_addr, a_base + 8 * (rec & 7u), desc, idesc, acc);
                            acc = 1;
                            desc += desc_step;
                            rec = s_rec[++s];
                        } while (!(rec & 24u));
                    }
                    tc_commit(bar_free + 8 * stage);
                    tc_commit(bar_dready + 8 * dbuf);
                    ++q_flush;
                }
            }
            __syncwarp();
        }
        item = run_end;
    }
    if (!ok) atomicExch(flag, 2);

    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// B operand of one effect set: 256 bytes per step.  K index k = 4s + b contracts marker bit 8b + s of the
// step's word (operand byte value 2^s); the byte of (k, digit) sits at (k >> 4) * 128 + digit * 16 + (k & 15).
// Digits are those of round(delta * 2^(exp2 - s)); markers of the word outside the step's block stay zero.
void build_local_tc5_fragments(const std::vector<double>& delta /* [M] */, const std::vector<uint2>& steps,
                               const uint32_t* block_offsets, const uint32_t* block_markers, int exp2, std::vector<uint32_t>& frag) {
    frag.assign(steps.size() * 64, 0u);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(frag.data());
    for (size_t st = 0; st < steps.size(); ++st) {
        const uint32_t w = steps[st].x, blk = steps[st].y & 0x7fffffffu;
        const uint32_t m_lo = block_markers[block_offsets[blk]], m_hi = m_lo + (block_offsets[blk + 1] - block_offsets[blk]);
        for (uint32_t bit = 0; bit < 32; ++bit) {
            const uint32_t m = w * 32 + bit;
            if (m < m_lo || m >= m_hi) continue;
            const uint32_t b = bit >> 3, s = bit & 7, k = 4 * s + b;
            long long q = std::isfinite(delta[m]) ? std::llrint(std::ldexp(delta[m], exp2 - (int) s)) : 0;
            for (uint32_t dg = 0; dg < 8; ++dg) {
                const long long d8 = ((q + 128) & 255) - 128;
                q = (q - d8) >> 8;
                bytes[st * 256 + (k >> 4) * 128 + dg * 16 + (k & 15)] = (uint8_t) (int8_t) d8;
            }
        }
    }
}

// One pass over the genotypes for n_sets (<= kLocalTc5MaxSets) effect sets.  d_steps: the schedule with the
// first-step-of-block flag in bit 31 of .y; d_b / d_base / d_out / scale per set.  Asynchronous.
// Marker segments for a pass of n_sets effect sets: consecutive step ranges of at most `cap` steps (their B
// digits must fit shared memory) that end where a 128-byte line of the strand rows ends (32 marker words),
// so that every line of the genotypes is fetched from DRAM by one segment only; when not even 32 words'
// steps fit (blocks of a few markers each) a segment is cut at `cap` steps.  A block that a segment boundary
// cuts is summed in two parts (fp64 atomics on the output); whole blocks are stored plainly.  Cutting at block
// boundaries instead was measured and is slower: segments then start inside a line, which two segments fetch.
void local_tc5_segments(const std::vector<uint2>& steps, uint32_t n_sets, std::vector<uint32_t>& seg_off) {
    const uint32_t cap = std::min<uint32_t>(kMaxSegSteps, kBBytes / (256 * n_sets));
    const uint32_t n_steps = (uint32_t) steps.size();
    seg_off.assign(1, 0u);
    uint32_t pos = 0;
    while (pos < n_steps) {
        uint32_t best = 0, end = pos;
        for (uint32_t end_word = (steps[pos].x / 32 + 1) * 32;; end_word += 32) {
            while (end < n_steps && steps[end].x < end_word) ++end;
            if (end - pos > cap) break;
            best = end;
            if (end == n_steps) break;
        }
        if (best == 0) best = std::min(n_steps, pos + cap);
        seg_off.push_back(best);
        pos = best;
    }
}

int local_tc5_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t n_blocks, const uint2* d_steps,
                     const std::vector<uint32_t>& seg_off, const uint32_t* d_seg_off,
                     uint32_t n_sets, const uint32_t* const* d_b, const double* const* d_base, double* const* d_out, const double* scale) {
    GSB_REQUIRE(n_sets >= 1 && n_sets <= kLocalTc5MaxSets, GSB200_ERR_ARG, "local GEBV (tcgen05 path): %u effect sets in one pass", n_sets);
    LocalTc5Params P;
    memset(&P, 0, sizeof P);
    for (uint32_t q = 0; q < n_sets; ++q) { P.b[q] = (const uint4*) d_b[q]; P.base[q] = d_base[q]; P.out[q] = d_out[q]; P.scale[q] = scale[q]; }
    const uint32_t cap = std::min<uint32_t>(kMaxSegSteps, kBBytes / (256 * n_sets));
    const uint32_t n_seg = (uint32_t) seg_off.size() - 1;
    const uint32_t n_tiles = (n + 63) / 64;
    const uint64_t n_items64 = (uint64_t) n_seg * n_tiles;
    GSB_REQUIRE(n_items64 < (1ull << 31), GSB200_ERR_ARG, "local GEBV: too many work items");
    const uint32_t n_items = (uint32_t) n_items64;
    const uint32_t ctas = std::min<uint32_t>((uint32_t) ctx->sm_count, n_items);
    const uint32_t items_per_cta = (n_items + ctas - 1) / ctas;
    const uint32_t grid = (n_items + items_per_cta - 1) / items_per_cta;
    uint32_t seg_steps = 0;
    for (uint32_t k = 0; k < n_seg; ++k) seg_steps = std::max(seg_steps, seg_off[k + 1] - seg_off[k]);
    GSB_REQUIRE(seg_steps <= cap, GSB200_ERR_ARG, "local GEBV (tcgen05 path): internal: a segment of %u steps exceeds %u", seg_steps, cap);
    const uint32_t smem = seg_steps * 256 * n_sets + 256;
    GSB_CUDA_TRY(cudaFuncSetAttribute(local_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kBBytes + 256));
    // GSB200_LOCAL_TRACE=<file>: clock64 stamps of CTA 0's first stages (profiles/scripts/local_trace.py reads them)
    static const char* trace_path = getenv("GSB200_LOCAL_TRACE");
    static const uint32_t dbg = getenv("GSB200_LOCAL_DEBUG") ? (uint32_t) atoi(getenv("GSB200_LOCAL_DEBUG")) : 0u;
    long long* d_trace = nullptr;
    if (trace_path) {
        GSB_CUDA_TRY(cudaMalloc((void**) &d_trace, 9 * kTraceLen * sizeof(long long)));
        GSB_CUDA_TRY(cudaMemsetAsync(d_trace, 0, 9 * kTraceLen * sizeof(long long), ctx->stream));
    }
    local_tc5_kernel<<<grid, kThreads, smem, ctx->stream>>>(ctx->d_store, ctx->row_words(), ctx->plane_words, d_rows, n, n_blocks, d_steps,
                                                           (uint32_t) seg_off.back(), d_seg_off, n_tiles, n_items, items_per_cta, n_sets, P, ctx->d_flag + 2, dbg, d_trace);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    if (d_trace) {
        std::vector<long long> h(9 * kTraceLen);
        GSB_CUDA_TRY(cudaMemcpyAsync(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_trace);
        if (FILE* f = fopen(trace_path, "w")) {
            fprintf(f, "# stage/flush: expander warp 0 (enter, stage free, first store issued, stores done) | issuer (wants stage, has stage) | epilogue warp 0 (wants sums, has sums, done); cycles since the first stamp\n");
            long long t0 = h[0];
            for (uint32_t k = 0; k < kTraceLen; ++k) {
                fprintf(f, "%3u", k);
                for (uint32_t r = 0; r < 9; ++r) fprintf(f, " %9lld", h[r * kTraceLen + k] ? h[r * kTraceLen + k] - t0 : -1ll);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return GSB200_OK;
}

}  // namespace gsb
