// gebv_umma.cu — GEBV digit contraction on the 5th-generation tensor cores (tcgen05, TMEM).
//
// Same arithmetic as gebv_digits_kernel (evaluate.cu): BV_i = base + sum_m delta_m * (bit0 + bit1),
// delta as 8 balanced base-256 digits = the N = 8 columns of B, strand bits expanded to weighted
// 0 / 2^s bytes = A, int32 accumulation, digit sums recombined in fp64 afterwards.  What changes
// is where the operands live.  The legacy mma.sync path issues one IMMA per 16 x 32 x 8 tile from
// every warp, and that issue traffic shares the schedulers with the LOP3 expansion (the two did
// not overlap: 37 cycles per 32 markers x 16 individuals).  Here:
//   * a CTA of 128 threads owns 128 strand rows (64 individuals); thread r owns row r = TMEM lane r;
//   * per 128 markers a thread makes its 32 operand registers (one LOP3 each) and writes them with
//     ONE tcgen05.st.32x32b.x32 into its lane of a TMEM A stage — A never touches shared memory,
//     whose 128 B/clk could not carry one byte per marker and strand at the HBM rate;
//   * one elected thread issues tcgen05.mma.kind::i8 (M = 128, N = 8, K = 32) with A from TMEM and
//     B from shared memory; accumulators stay in TMEM until the segment ends (tcgen05.ld);
//   * three TMEM A stages rotate under tcgen05.commit -> mbarrier; the last of the CTA's four warps to
//     have written a stage issues its four mma; packed rows and B fragments arrive by 16-byte
//     cp.async, two chunks ahead (B has two more buffers: its readers are asynchronous);
//   * persistent CTAs (TMEM allocated once) pull tasks from an atomic counter.
//
// STATUS: EXPERIMENTAL, opt-in with GSB200_GEBV_PATH=umma.  Results are identical to the mma.sync
// path (tests/test_gpu_parity.py::test_gebv_tcgen05_path), but on the wheat shape it is slower:
// 0.50-0.73 ms per 100 k individuals across the variants tried (dedicated issuer warp with
// per-thread or per-warp arrivals, persistent or not, 1-3 chunks of prefetch) against 0.335 ms for
// gebv_digits_kernel.  With the expansion AND the TMEM stores removed the pipeline alone still
// took 0.45 ms, and with no mma issued 0.48 ms: the per-128-marker handshake (stage hand-off,
// commit -> mbarrier round trip) bounds it, not the ALU work this design was meant to free.
// Round-2 lead: tiles of >= 512 markers per hand-off (needs cta_group::2 or a 256-column stage).
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

using namespace gsb;

namespace gsb {

namespace {

constexpr uint32_t kProducers = 128;          // one thread per strand row = TMEM lane
constexpr uint32_t kThreads = kProducers;     // no issuer warp: the last row to arrive issues the mma
constexpr uint32_t kChunkWords = 16;          // 512 markers = 4 tiles of 128 per packed-row chunk
constexpr uint32_t kRowStride = 20;           // words: 64 B + 16 B pad
constexpr uint32_t kPackedWords = 128 * kRowStride;
constexpr uint32_t kBChunkWords = 16 * 64;    // 16 k-steps x 256 B
constexpr uint32_t kPackedBufs = 3;           // cp.async runs kPackedBufs - 1 chunks ahead of the expansion
constexpr uint32_t kBBufs = kPackedBufs + 2;   // B is read asynchronously by the mma: two more chunks of slack
constexpr uint32_t kSmemBytes = (kPackedBufs * kPackedWords + kBBufs * kBChunkWords) * 4;
constexpr uint32_t kAStages = 3;
constexpr uint32_t kTmemCols = 128;           // D: 8 columns, A stages: 3 x 32 columns
constexpr uint32_t kSpinLimit = 1u << 20;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded waits: a protocol bug must end in a flagged wrong answer, never in a hung GPU
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
// the common case (phase already complete) costs one try_wait; only a miss enters the bounded loop
__device__ __noinline__ bool mbar_wait_slow(uint32_t addr, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity, bool& ok) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done) ok = mbar_wait_slow(addr, parity) && ok;
}
// arrives on `bar` once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive(uint32_t addr) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(addr) : "memory");
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
    return d;                                         // base offset 0, layout type 0 = no swizzle
}

constexpr uint32_t kIdesc = (2u << 4)      // D: s32
                          | (0u << 7)      // A: u8
                          | (1u << 10)     // B: s8
                          | (1u << 17)     // N >> 3
                          | (8u << 24);    // M >> 4

// Persistent: a CTA allocates its TMEM once and pulls (64 individuals x marker segment) tasks from
// an atomic counter; the mbarrier phases simply keep counting across tasks.
__global__ void __launch_bounds__(kThreads) gebv_umma_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ rows, uint32_t n, const uint32_t* __restrict__ bfrag /* 64 words per marker word */,
        uint32_t n_segments, uint32_t seg_chunks, unsigned long long* __restrict__ digit_sums, int* __restrict__ flag,
        unsigned int* __restrict__ task_counter) {
    extern __shared__ __align__(128) uint32_t s_dyn[];
    uint32_t (*s_b)[kBChunkWords] = reinterpret_cast<uint32_t (*)[kBChunkWords]>(s_dyn);
    uint32_t (*s_packed)[kPackedWords] = reinterpret_cast<uint32_t (*)[kPackedWords]>(s_dyn + kBBufs * kBChunkWords);
    __shared__ __align__(8) uint64_t s_packed_full[kPackedBufs], s_packed_empty[kPackedBufs];   // cp.async landed / rows consumed
    __shared__ __align__(8) uint64_t s_a_empty[kAStages];                   // operand stage read by the mma (tcgen05.commit)
    __shared__ uint32_t s_a_count[kAStages];                               // rows that have written the stage
    __shared__ uint32_t s_tmem, s_task;

    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t n_groups = (n + 63) / 64, n_tasks = n_groups * n_segments;
    const uint32_t total_chunks = plane_words / kChunkWords;

    if (tid == 0) {
        // one arrival per WARP: 128 single-thread arrivals on one barrier serialise in the shared-memory
        // atomic unit and alone cost more than the whole tile should
        for (uint32_t i = 0; i < kPackedBufs; ++i) { mbar_init(&s_packed_full[i], kProducers / 32); mbar_init(&s_packed_empty[i], kProducers / 32); }
        for (uint32_t i = 0; i < kAStages; ++i) { mbar_init(&s_a_empty[i], 1); s_a_count[i] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    bool ok = true;
    const uint32_t bar_packed_full = smem_u32(s_packed_full), bar_packed_empty = smem_u32(s_packed_empty);
    const uint32_t bar_a_empty = smem_u32(s_a_empty), cnt_a = smem_u32(s_a_count);
    const uint32_t b_base = smem_u32(s_dyn);

    // running counts across tasks: they define every barrier's phase parity
    uint32_t tiles_done = 0, chunks_done = 0, tasks_done = 0;
    uint32_t st_stage = 0, st_phase = 0;                   // next operand stage and the parity of its use count

    for (;;) {
        tc_fence_before();
        __syncthreads();                                   // everyone has left the previous task (its D was read)
        if (tid == 0) s_task = atomicAdd(task_counter, 1u);
        __syncthreads();
        tc_fence_after();
        const uint32_t task = s_task;
        if (task >= n_tasks) break;
        const uint32_t group = task % n_groups, seg = task / n_groups;    // neighbours share B
        const uint32_t ch_lo = seg * seg_chunks, ch_hi = min(ch_lo + seg_chunks, total_chunks);
        const uint32_t n_chunks = ch_hi > ch_lo ? ch_hi - ch_lo : 0, n_tiles = 4 * n_chunks;

        {
            // ---------------- producers: thread r owns strand row r
            // staging role: copy `it` moves 16 bytes of strand row it*32 + tid/4 (row = 2*individual + strand)
            const uint32_t* stage_src[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t r = it * 32 + (tid >> 2);
                const uint32_t ind = min(group * 64 + (r >> 1), n - 1);
                stage_src[it] = store + (uint64_t) rows[ind] * row_words + (r & 1) * plane_words + (tid & 3) * 4;
            }
            auto stage = [&](uint32_t c) {                 // c: chunk of this task; cc: running chunk count
                const uint32_t cc = chunks_done + c, pb = cc % kPackedBufs;
                // packed buffer pb was last read as running chunk cc - kPackedBufs: wait until all rows were
                // consumed.  B buffer cc % kBBufs was last read by the mma of running chunk cc - kBBufs,
                // complete by now (this thread has waited for a_empty of a later tile, or for s_done).
                if (cc >= kPackedBufs) mbar_wait(bar_packed_empty + 8 * pb, ((cc - kPackedBufs) / kPackedBufs) & 1, ok);
                uint32_t* packed = s_packed[pb];
#pragma unroll
                for (int it = 0; it < 4; ++it)
                    cp_async16(packed + (it * 32 + (tid >> 2)) * kRowStride + (tid & 3) * 4, stage_src[it] + (size_t) (ch_lo + c) * kChunkWords);
                // B of the chunk: 16 k-steps x 256 B = 4 KB = 256 pieces of 16 B, two per thread
                uint32_t* b = s_b[cc % kBBufs];
                const uint32_t* src = bfrag + (size_t) (ch_lo + c) * kBChunkWords;
                cp_async16(b + tid * 4, src + tid * 4);
                cp_async16(b + (tid + 128) * 4, src + (tid + 128) * 4);
            };

            // the accumulators start from zero (every mma accumulates); the store is covered by this
            // row's tcgen05.wait::st before it publishes its first tile
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};"
                         :: "r"(tmem + ((warp * 32) << 16)), "r"(0u) : "memory");
            // one cp.async group per chunk slot, committed even when empty, so that "all but the newest
            // kPackedBufs - 1 groups have landed" always means "chunk c has landed"
            for (uint32_t c = 0; c < kPackedBufs - 1; ++c) { if (c < n_chunks) stage(c); cp_async_commit(); }
            bool pending = false;                          // a tcgen05.st of the previous tile is in flight
            uint32_t pending_stage = 0, pending_tile = 0, tile = 0;
            // A warp's operand registers are in TMEM: count it in; the LAST of the 4 warps to arrive issues
            // the four tcgen05.mma of the tile (no thread ever sleeps waiting for "all rows written").
            auto publish = [&](uint32_t stg, uint32_t tl) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if ((tid & 31) == 0) {
                    uint32_t old;
                    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt_a + 4 * stg) : "memory");
                    if (old == kProducers / 32 - 1) {
                        asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" :: "r"(cnt_a + 4 * stg), "r"(0u) : "memory");
                        tc_fence_after();
                        const uint32_t cc = chunks_done + (tl >> 2);
                        const uint32_t b_addr = b_base + (cc % kBBufs) * (kBChunkWords * 4) + (tl & 3) * 1024, a_col = 8 + 32 * stg;
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)      // always accumulating: tiles may be issued by different threads
                            umma_i8_ts(tmem, tmem + a_col + 8 * k, b_descriptor(b_addr + k * 256), kIdesc, 1u);
                        tc_commit(bar_a_empty + 8 * stg);
                    }
                }
            };
            for (uint32_t c = 0; c < n_chunks; ++c) {
                const uint32_t cc = chunks_done + c;
                if (c + kPackedBufs - 1 < n_chunks) stage(c + kPackedBufs - 1);
                cp_async_commit();
                const uint32_t pb = cc % kPackedBufs;
                cp_async_wait<kPackedBufs - 1>();                            // this thread's pieces of chunk c
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar_packed_full + 8 * pb);  // ... this warp's
                mbar_wait(bar_packed_full + 8 * pb, (cc / kPackedBufs) & 1, ok);   // ... everyone's
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // B bytes -> visible to the tensor core's proxy
                const uint32_t* packed = s_packed[pb] + tid * kRowStride;
#pragma unroll 1
                for (uint32_t t = 0; t < 4; ++t, ++tile) {
                    const uint4 x = *reinterpret_cast<const uint4*>(packed + t * 4);
                    const uint32_t w[4] = { x.x, x.y, x.z, x.w };
                    uint32_t r[32];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int j = 0; j < 8; ++j) r[k * 8 + j] = w[k] & (0x01010101u << j);   // byte i = bit 8i + j, weight 2^j
                    if (pending) publish(pending_stage, pending_tile);   // its store has had this tile's expansion to complete
                    if (tiles_done + tile >= kAStages) {                     // the mma that read this stage kAStages tiles ago
                        mbar_wait(bar_a_empty + 8 * st_stage, st_phase ^ 1, ok);
                        tc_fence_after();
                    }
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                        :: "r"(tmem + ((warp * 32) << 16) + 8 + 32 * st_stage),
                           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                        : "memory");
                    pending = true;
                    pending_stage = st_stage;
                    pending_tile = tile;
                    if (++st_stage == kAStages) { st_stage = 0; st_phase ^= 1; }
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar_packed_empty + 8 * pb);
            }
            if (pending) publish(pending_stage, pending_tile);
            // every tile's mma has completed once each operand stage's last commit has arrived
            {
                const uint32_t total = tiles_done + n_tiles;
#pragma unroll
                for (uint32_t sg = 0; sg < kAStages; ++sg) {
                    const uint32_t uses = total > sg ? (total - sg + kAStages - 1) / kAStages : 0;
                    if (uses) mbar_wait(bar_a_empty + 8 * sg, (uses - 1) & 1, ok);
                }
            }
            tc_fence_after();

            // epilogue: lane r of TMEM = strand row r; its 8 digit sums are columns 0..7
            uint32_t d[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
                         : "r"(tmem + ((warp * 32) << 16)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const uint32_t ind = group * 64 + (tid >> 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int both = (int) d[j] + __shfl_xor_sync(0xffffffffu, (int) d[j], 1);      // the individual's two strands
                if ((tid & 1) == 0 && ind < n && n_tiles > 0) atomicAdd(&digit_sums[(size_t) ind * 8 + j], (unsigned long long) (long long) both);
            }
        }
        tiles_done += n_tiles;
        chunks_done += n_chunks;
        ++tasks_done;
    }
    if (!ok) atomicExch(flag, 2);

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// B operand for the kernel above: per marker word 256 bytes, k-step-major: byte of (k, digit) at
// (k >> 4) * 128 + digit * 16 + (k & 15), with k = 4 * j + i <-> marker bit 8 * i + j of the word.
void build_umma_fragments(const std::vector<double>& delta, uint32_t n_markers, uint32_t plane_words, int exp2,
                          std::vector<uint32_t>& frag) {
    frag.assign((size_t) plane_words * 64, 0u);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(frag.data());
    for (uint32_t m = 0; m < n_markers; ++m) {
        const uint32_t w = m >> 5, bit = m & 31, i = bit >> 3, j = bit & 7, k = 4 * j + i;
        long long q = std::isfinite(delta[m]) ? std::llrint(std::ldexp(delta[m], exp2 - (int) j)) : 0;
        for (uint32_t dg = 0; dg < 8; ++dg) {
            const long long d8 = ((q + 128) & 255) - 128;
            q = (q - d8) >> 8;
            bytes[(size_t) w * 256 + (k >> 4) * 128 + dg * 16 + (k & 15)] = (uint8_t) (int8_t) d8;
        }
    }
}

int gebv_umma_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, const uint32_t* d_frag, unsigned long long* d_digits) {
    const uint32_t chunks = ctx->plane_words / kChunkWords, n_groups = (n + 63) / 64;
    const uint32_t ctas = (uint32_t) ctx->sm_count * 4;          // 4 x 128 TMEM columns, 4 x 50 KB of shared memory per SM
    // about 8 tasks per CTA for balance; at most 32768 markers per task (int32 sums)
    uint32_t n_segments = (uint32_t) std::min<uint64_t>(chunks, ((uint64_t) ctas * 8 + n_groups - 1) / n_groups);
    n_segments = std::max<uint32_t>(std::max<uint32_t>(n_segments, 1), (chunks + 63) / 64);
    const uint32_t seg_chunks = (chunks + n_segments - 1) / n_segments;
    n_segments = (chunks + seg_chunks - 1) / seg_chunks;
    unsigned int* d_counter = reinterpret_cast<unsigned int*>(ctx->d_flag) + 1;
    GSB_CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), ctx->stream));
    const uint32_t grid = std::min<uint32_t>(ctas, n_groups * n_segments);
    GSB_CUDA_TRY(cudaFuncSetAttribute(gebv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBytes));
    gebv_umma_kernel<<<grid, kThreads, kSmemBytes, ctx->stream>>>(
        ctx->d_store, ctx->row_words(), ctx->plane_words, d_rows, n, d_frag, n_segments, seg_chunks, d_digits, ctx->d_flag, d_counter);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB200_OK;
}

}  // namespace gsb
