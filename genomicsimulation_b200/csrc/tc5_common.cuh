// tc5_common.cuh — the handful of sm_100a PTX wrappers the two tcgen05 kernels share
// (gebv_tc5.cu, local_tc5.cu): mbarriers with BOUNDED waits, tcgen05 fences / commit, the
// kind::i8 MMA with A in TMEM and B behind a shared-memory descriptor.
#pragma once

#include <stdint.h>

namespace gsb {
namespace tc5 {

constexpr uint32_t kSpinLimit = 1u << 22;

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
// Once a wait has timed out the rest are skipped: the kernel runs to its end with a wrong answer and
// the flag set — a protocol bug ends in a reported error, never in a hung GPU.
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

// D[tmem] (+)= A[tmem] * B[smem descriptor], u8 x s8 -> s32, M = 128, K = 32; N comes with the instruction descriptor
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// K-major, no swizzle: core matrices of 8 rows (N) x 16 bytes (K); LBO = distance between the two core
// matrices a K = 32 step spans, SBO = distance between 8-row groups along N.
__device__ __forceinline__ uint64_t b_descriptor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t) ((128u >> 4) & 0x3fff) << 16;     // leading byte offset
    d |= (uint64_t) ((256u >> 4) & 0x3fff) << 32;     // stride byte offset
    d |= (uint64_t) 1 << 46;                          // descriptor version (sm_100)
    return d;
}

// instruction descriptor: D s32, A u8, B s8, M = 128, N = n_cols (a multiple of 8)
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t n_cols) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((n_cols >> 3) << 17) | (8u << 24);
}

}  // namespace tc5
}  // namespace gsb
