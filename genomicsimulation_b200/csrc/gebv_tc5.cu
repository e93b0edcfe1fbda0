// gebv_tc5.cu — GEBV digit contraction on the 5th-generation tensor cores, second design.
//
// Arithmetic as in gebv_counts_kernel (evaluate.cu), per strand: BV_i = base + sum_m delta_m * (bit0 + bit1),
// delta as 8 balanced base-256 digits = the N = 8 columns of B, strand bits expanded to weighted
// 0 / 2^s bytes = A, int32 accumulation, digit sums recombined in fp64 afterwards.  The legacy
// mma.sync path is bounded by the IMMA issue rate (11 cycles per m16n8k32 and scheduler:
// profiles/r1_pipe_overlap.txt); tcgen05.mma.kind::i8 does the same 128 x 8 x 32 tile in 4.
//
//   * a CTA = 4 expander warps + 1 issuer warp owns 128 strand rows (64 individuals) at a time;
//   * expander lane (g, c) loads 16 bytes (words 4c..4c+3 of a 16-word chunk) of four of its
//     warp's 32 rows straight into registers — the 4 lanes of a row cover 64 contiguous bytes —
//     four chunks ahead; no shared-memory staging of genotypes at all;
//   * an operand register is `word & immediate`: byte b = bit 8b + s of the word at weight 2^s.
//     tcgen05.st.16x256b puts registers {0,1} of lane (g, c) into TMEM lane g, columns 2c, 2c+1 and
//     {2,3} into lane g + 8 (the mma.sync C-fragment layout; profiles/scripts/sttm_layout_probe.cu),
//     so K-step (i, p) takes k bytes 8c..8c+3 from bits 8b + 2p and 8c+4..8c+7 from bits 8b + 2p + 1
//     of word 4c + i.  One .x8 store = 8 K-steps = 256 markers of 16 rows;
//   * the B digits of the CTA's marker segment (<= 16 chunks = 64 KB) stay in shared memory for
//     all the row groups the CTA processes: no per-stage B traffic, no per-stage proxy fence;
//   * three TMEM operand stages of 256 markers; lane 0 of the issuer warp waits for a stage's four
//     warps, issues 8 tcgen05.mma (M = 128, N = 8, K = 32, A from TMEM, B from shared memory) and
//     commits the stage back; accumulators are double-buffered so the digit sums of one row group
//     are read (tcgen05.ld) and added to global memory while the next group is being expanded.
//
// Every mbarrier wait is bounded: a protocol bug ends in a flagged error, never in a hung GPU.
//
// STATUS: EXPERIMENTAL, opt-in with GSB200_GEBV_PATH=tc5.  Bit-identical to gebv_counts_kernel's
// answer (tests/test_gpu_parity.py::test_gebv_tcgen05_path) and 2.3x slower on the wheat shape
// (0.58 ms against 0.25 ms per 100 k individuals).  A per-stage clock64 trace
// (profiles/r1_tc5_timeline.txt) shows why: the one-bit-to-one-byte expansion makes the A operand
// 8x the packed genotype, and that operand has to go through TMEM twice.  One tcgen05.st.16x256b.x8
// (4 KB) occupies its warp for ~200 cycles, so a CTA's four warps deliver a 256-marker stage in
// ~400 cycles; the eight tcgen05.mma (M = 128, N = 8) that read the stage back take the issuing
// thread ~700 cycles, ~87 each, for a nominal 4 cycles of math: at N = 8 the instruction is bound by
// streaming its 4 KB A operand out of TMEM.  Both sit at ~1000 cycles per stage and CTA, i.e.
// ~64 bytes of operand per cycle and SM, while the HBM rate would need 184 (8 x 23 bytes).  With
// the mma removed the kernel takes 0.56 ms, with the stores and the loads removed 0.61 ms: either
// side alone is as slow as the whole.  The tensor cores' operand paths are sized for operands that
// are reused across N = 128..256 columns; a contraction with N = 8 and no reuse cannot feed them,
// so the product path stays on mma.sync, where A registers come straight from the register file.
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

using namespace gsb;

namespace gsb {

namespace {

constexpr uint32_t kExpanders = 128;          // 4 warps; warp w owns TMEM lanes 32w .. 32w + 31
constexpr uint32_t kThreads = kExpanders + 32;
constexpr uint32_t kChunkWords = 16;          // 512 markers = 64 bytes per strand row = 2 operand stages
constexpr uint32_t kMaxSegChunks = 16;        // B of a segment: 16 x 4 KB of shared memory
constexpr uint32_t kStages = 3;
constexpr uint32_t kStageCols = 64;           // 8 K-steps x 8 columns
constexpr uint32_t kACol0 = 32;               // columns 0..15: two accumulators
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kDepth = 4;                // chunks of row data in flight per lane
constexpr uint32_t kSpinLimit = 1u << 20;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t addr, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
// once a wait has timed out the rest are skipped: the kernel runs to its end with a wrong answer and the flag set
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity, bool& ok) {
    if (!ok) return;
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done) ok = mbar_wait_slow(addr, parity);
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(addr) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem descriptor], u8 x s8 -> s32, M = 128, N = 8, K = 32
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// K-major, no swizzle: core matrices of 8 rows x 16 bytes; LBO = distance between the two core
// matrices a K = 32 step spans, SBO = distance between 8-row groups (one group: N = 8).
__device__ __forceinline__ uint64_t b_descriptor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t) ((128u >> 4) & 0x3fff) << 16;     // leading byte offset
    d |= (uint64_t) ((256u >> 4) & 0x3fff) << 32;     // stride byte offset
    d |= (uint64_t) 1 << 46;                          // descriptor version (sm_100)
    return d;
}

constexpr uint32_t kIdesc = (2u << 4)      // D: s32
                          | (0u << 7)      // A: u8
                          | (1u << 10)     // B: s8
                          | (1u << 17)     // N >> 3
                          | (8u << 24);    // M >> 4

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t pick(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// 256 markers of rows (lane g: wa, lane g + 8: wb) -> 8 K-steps of operand bytes
__device__ __forceinline__ void store_stage(uint32_t taddr, uint32_t wa0, uint32_t wa1, uint32_t wb0, uint32_t wb1) {
    uint32_t r[32];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t wa = t < 4 ? wa0 : wa1, wb = t < 4 ? wb0 : wb1;
        const uint32_t m_lo = 0x01010101u << (2 * (t & 3)), m_hi = 0x02020202u << (2 * (t & 3));
        r[4 * t + 0] = wa & m_lo;
        r[4 * t + 1] = wa & m_hi;
        r[4 * t + 2] = wb & m_lo;
        r[4 * t + 3] = wb & m_hi;
    }
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// Work item = (marker segment, group of 64 individuals), segment-major; CTA b owns items
// [b * items_per_cta, (b + 1) * items_per_cta) and reloads B when the segment changes.
__global__ void __launch_bounds__(kThreads, 2) gebv_tc5_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ rows, uint32_t n, const uint4* __restrict__ bfrag /* 256 uint4 per chunk */,
        uint32_t n_groups, uint32_t n_items, uint32_t items_per_cta, uint32_t seg_chunks,
        unsigned long long* __restrict__ digit_sums, int* __restrict__ flag) {
    extern __shared__ __align__(128) uint4 s_b[];                   // seg_chunks x 4 KB
    __shared__ __align__(8) uint64_t s_full[kStages], s_free[kStages], s_dready[2], s_dfree[2];
    __shared__ uint32_t s_tmem;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t total_chunks = plane_words / kChunkWords;
    if (tid == 0) {
        for (uint32_t i = 0; i < kStages; ++i) { mbar_init(&s_full[i], kExpanders / 32); mbar_init(&s_free[i], 1); }
        for (uint32_t i = 0; i < 2; ++i) { mbar_init(&s_dready[i], 1); mbar_init(&s_dfree[i], kExpanders / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
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
    bool ok = true;

    const uint32_t item_lo = min(blockIdx.x * items_per_cta, n_items), item_hi = min(item_lo + items_per_cta, n_items);
    // running counts across the CTA's items: they define every barrier's phase parity
    uint32_t stage = 0, stage_use = 0;       // next operand stage and how often it has been used
    uint32_t items_done = 0;
    bool pend = false;                       // expanders: the previous item's digit sums are still in TMEM
    uint32_t pend_group = 0, pend_item = 0;

    const uint32_t g = lane >> 2, c = lane & 3;
    // expanders: read one item's accumulator (TMEM lane = strand row = 2 * individual + strand)
    auto epilogue = [&]() {
        const uint32_t dbuf = pend_item & 1;
        mbar_wait(bar_dready + 8 * dbuf, (pend_item >> 1) & 1, ok);
        tc_fence_after();
        uint32_t d[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
                     : "r"(tmem + ((warp * 32) << 16) + 8 * dbuf) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dfree + 8 * dbuf);
        const uint32_t ind = pend_group * 64 + (tid >> 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int both = (int) d[j] + __shfl_xor_sync(0xffffffffu, (int) d[j], 1);      // the individual's two strands
            if ((tid & 1) == 0 && ind < n) atomicAdd(&digit_sums[(size_t) ind * 8 + j], (unsigned long long) (long long) both);
        }
        pend = false;
    };

    for (uint32_t item = item_lo; item < item_hi;) {
        const uint32_t seg = item / n_groups, run_end = min(item_hi, (seg + 1) * n_groups);
        const uint32_t ch_lo = seg * seg_chunks, segc = min(seg_chunks, total_chunks - ch_lo);
        const uint32_t group_lo = item - seg * n_groups, n_run = run_end - item;

        // ---- B of this segment -> shared memory (the previous segment's mma have all completed)
        if (warp < 4 && pend) epilogue();
        tc_fence_before();
        __syncthreads();
        for (uint32_t i = tid; i < segc * 256; i += kThreads) s_b[i] = __ldg(bfrag + (size_t) ch_lo * 256 + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to the tensor core's proxy
        __syncthreads();
        tc_fence_after();

        if (warp < 4) {
            // ---------------- expanders
            // lane (g, c) of warp w holds rows 32w + 16h + g + 8e (h, e = 0, 1): buffer index 2h + e
            auto row_ptr = [&](uint32_t group, uint32_t q) -> const uint4* {
                const uint32_t r = warp * 32 + (q >> 1) * 16 + g + (q & 1) * 8;
                const uint32_t ind = min(min(group, n_groups - 1) * 64 + (r >> 1), n - 1);
                return reinterpret_cast<const uint4*>(store + (uint64_t) rows[ind] * row_words + (r & 1) * plane_words + ch_lo * kChunkWords) + c;
            };
            const uint4* lptr[4];
            const uint4* nptr[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { lptr[q] = row_ptr(group_lo, q); nptr[q] = row_ptr(group_lo + 1, q); }
            uint32_t lgroup = group_lo, lch = 0;               // load side: next chunk to request
            uint4 buf[kDepth][4];
            auto request = [&](uint4 (&b)[4]) {
#pragma unroll
                for (int q = 0; q < 4; ++q) b[q] = ldg_stream(lptr[q] + lch * 4);
                if (++lch == segc) {
                    lch = 0;
                    ++lgroup;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { lptr[q] = nptr[q]; nptr[q] = row_ptr(lgroup + 1, q); }
                }
            };
            const uint32_t total = n_run * segc;
#pragma unroll
            for (int k = 0; k < (int) kDepth; ++k) if ((uint32_t) k < total) request(buf[k]);

            uint32_t cgroup = group_lo, cch = 0;               // consume side
            const uint32_t epi_at = segc > 1 ? 1u : 0u;
            auto consume = [&](const uint4 (&b)[4]) {
                if (pend && cch == epi_at) epilogue();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (stage_use > 0) {                       // the mma that read this stage kStages stages ago
                        mbar_wait(bar_free + 8 * stage, (stage_use - 1) & 1, ok);
                        tc_fence_after();
                    }
                    const uint32_t taddr = tmem + ((warp * 32) << 16) + kACol0 + kStageCols * stage;
                    store_stage(taddr, pick(b[0], 2 * half), pick(b[0], 2 * half + 1), pick(b[1], 2 * half), pick(b[1], 2 * half + 1));
                    store_stage(taddr + (16u << 16), pick(b[2], 2 * half), pick(b[2], 2 * half + 1), pick(b[3], 2 * half), pick(b[3], 2 * half + 1));
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * stage);
                    if (++stage == kStages) { stage = 0; ++stage_use; }
                }
                if (++cch == segc) {
                    cch = 0;
                    pend = true; pend_group = cgroup; pend_item = items_done;
                    ++cgroup; ++items_done;
                }
            };
            for (uint32_t cf = 0; cf < total; cf += kDepth) {
#pragma unroll
                for (int k = 0; k < (int) kDepth; ++k) {
                    if (cf + k < total) {
                        consume(buf[k]);
                        if (cf + k + kDepth < total) request(buf[k]);
                    }
                }
            }
        } else {
            // ---------------- issuer
            if (lane == 0) {
                for (uint32_t it = 0; it < n_run; ++it, ++items_done) {
                    const uint32_t dbuf = items_done & 1;
                    if (items_done >= 2) {                     // the expanders have read this accumulator's previous sums
                        mbar_wait(bar_dfree + 8 * dbuf, ((items_done >> 1) - 1) & 1, ok);
                        tc_fence_after();
                    }
                    for (uint32_t s = 0; s < 2 * segc; ++s) {
                        mbar_wait(bar_full + 8 * stage, stage_use & 1, ok);
                        tc_fence_after();
                        const uint32_t a_col = kACol0 + kStageCols * stage, b_addr = b_base + s * 2048;
#pragma unroll
                        for (uint32_t t = 0; t < 8; ++t)
                            umma_i8_ts(tmem + 8 * dbuf, tmem + a_col + 8 * t, b_descriptor(b_addr + t * 256), kIdesc, (s | t) != 0);
                        tc_commit(bar_free + 8 * stage);
                        if (++stage == kStages) { stage = 0; ++stage_use; }
                    }
                    tc_commit(bar_dready + 8 * dbuf);
                }
            }
            __syncwarp();
        }
        item = run_end;
    }
    if (warp < 4 && pend) epilogue();
    if (!ok) atomicExch(flag, 2);

    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// B operand for the kernel above: 256 bytes per K-step, 16 K-steps per 16-word chunk.  K-step
// i * 4 + p of a chunk contracts, for c = 0..3 and b = 0..3, k = 8c + b with bit 8b + 2p of word
// 4c + i and k = 8c + 4 + b with bit 8b + 2p + 1; the byte of (k, digit) sits at
// (k >> 4) * 128 + digit * 16 + (k & 15).  Digits are those of round(delta * 2^(exp2 - s)).
void build_tc5_fragments(const std::vector<double>& delta, uint32_t n_markers, uint32_t plane_words, int exp2,
                         std::vector<uint32_t>& frag) {
    frag.assign((size_t) plane_words * 64, 0u);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(frag.data());
    for (uint32_t m = 0; m < n_markers; ++m) {
        const uint32_t w = m >> 5, bit = m & 31, b = bit >> 3, s = bit & 7, p = s >> 1, hi = s & 1;
        const uint32_t chunk = w / kChunkWords, wi = w % kChunkWords, c = wi >> 2, i = wi & 3;
        const uint32_t k = 8 * c + 4 * hi + b, step = chunk * 16 + i * 4 + p;
        long long q = std::isfinite(delta[m]) ? std::llrint(std::ldexp(delta[m], exp2 - (int) s)) : 0;
        for (uint32_t dg = 0; dg < 8; ++dg) {
            const long long d8 = ((q + 128) & 255) - 128;
            q = (q - d8) >> 8;
            bytes[(size_t) step * 256 + (k >> 4) * 128 + dg * 16 + (k & 15)] = (uint8_t) (int8_t) d8;
        }
    }
}

int gebv_tc5_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, const uint32_t* d_frag, unsigned long long* d_digits) {
    const uint32_t chunks = ctx->plane_words / kChunkWords, n_groups = (n + 63) / 64;
    const uint32_t n_segments = (chunks + kMaxSegChunks - 1) / kMaxSegChunks;
    const uint32_t seg_chunks = (chunks + n_segments - 1) / n_segments;
    const uint64_t n_items64 = (uint64_t) n_groups * n_segments;
    GSB_REQUIRE(n_items64 < (1ull << 31), GSB200_ERR_ARG, "gebv: too many work items");
    const uint32_t n_items = (uint32_t) n_items64;
    const uint32_t ctas = std::min<uint32_t>((uint32_t) ctx->sm_count * 2, n_items);
    const uint32_t items_per_cta = (n_items + ctas - 1) / ctas;
    const uint32_t grid = (n_items + items_per_cta - 1) / items_per_cta;
    const uint32_t smem = seg_chunks * 4096;
    GSB_CUDA_TRY(cudaFuncSetAttribute(gebv_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kMaxSegChunks * 4096)));
    gebv_tc5_kernel<<<grid, kThreads, smem, ctx->stream>>>(
        ctx->d_store, ctx->row_words(), ctx->plane_words, d_rows, n, reinterpret_cast<const uint4*>(d_frag),
        n_groups, n_items, items_per_cta, seg_chunks, d_digits, ctx->d_flag + 2);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    // experimental path: check the protocol flag right away
    int flag = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&flag, ctx->d_flag + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        GSB_CUDA_TRY(cudaMemsetAsync(ctx->d_flag + 2, 0, sizeof(int), ctx->stream));
        GSB_REQUIRE(false, GSB200_ERR_CUDA, "gebv (tcgen05 path): a pipeline wait timed out");
    }
    return GSB200_OK;
}

}  // namespace gsb
