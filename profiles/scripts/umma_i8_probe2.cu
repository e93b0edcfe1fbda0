// umma_i8_probe2.cu — cycles per tcgen05.mma.kind::i8 (M = 128, K = 32) as a function of N, with the
// issuing loop reduced to the MMA itself (8 unrolled MMAs per trip, operands advanced by constant
// adds): A from TMEM (the expander path of local_tc5.cu) against A from shared memory.
// umma_i8_probe.cu had two runtime `%` per MMA in its loop and measured those (99 cycles whatever N).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I genomicsimulation_b200/csrc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc5_common.cuh"
using namespace gsb::tc5;

__device__ __forceinline__ void umma_i8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

template <int MODE>   // 0: A in TMEM, 1: A in shared memory, 2: A in TMEM, issued from a CONVERGED warp under elect.sync (operands provably uniform)
__global__ void __launch_bounds__(128, 1) probe(uint32_t n_cols, uint32_t b_step_bytes, long long* out) {
    extern __shared__ __align__(128) uint8_t s_b[];          // 64 KB of B + 32 KB of A
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 96 * 1024; i += 128) s_b[i] = (uint8_t) (i * 7u);
    if (tid == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (MODE == 2 && warp == 1) {
        const uint32_t idesc = idesc_i8(n_cols), bar = smem_u32(&s_bar), b_base = smem_u32(s_b);
        const uint64_t bdesc0 = b_descriptor(b_base);
        const uint64_t bstep = b_step_bytes >> 4;
        bool ok = true;
        const long long t0 = clock64();
#pragma unroll 1
        for (int trip = 0; trip < 64; ++trip) {
            uint64_t bd = bdesc0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t is_leader;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
                if (is_leader) umma_i8_ts(tmem, tmem + 256u + 8u * k, bd, idesc, 1u);
                __syncwarp();
                bd += bstep;
            }
        }
        const long long t1 = clock64();
        if ((tid & 31) == 0) {
            tc_commit(bar);
            mbar_wait(bar, 0, ok);
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0; out[2] = ok;
        }
    }
    if (MODE != 2 && tid == 0) {
        const uint32_t idesc = idesc_i8(n_cols), bar = smem_u32(&s_bar), b_base = smem_u32(s_b);
        const uint64_t bdesc0 = b_descriptor(b_base), adesc0 = b_descriptor(b_base + 65536u);
        const uint64_t bstep = b_step_bytes >> 4;
        bool ok = true;
        const long long t0 = clock64();
#pragma unroll 1
        for (int trip = 0; trip < 64; ++trip) {
            uint64_t bd = bdesc0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) umma_i8_ts(tmem, tmem + 256u + 8u * k, bd, idesc, 1u);
                else umma_i8_ss(tmem, adesc0 + (uint64_t) (k * 256), bd, idesc, 1u);      // 4 KB of A per K step
                bd += bstep;
            }
        }
        const long long t1 = clock64();
        tc_commit(bar);
        mbar_wait(bar, 0, ok);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0; out[2] = ok;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_out; long long h[3];
    cudaMalloc(&d_out, 24);
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    printf("# tcgen05.mma.kind::i8 M=128 K=32, B no-swizzle K-major in smem; 512 MMAs, 8 unrolled per trip; cycles per MMA (issue only | until complete)\n");
    const uint32_t ns[] = { 8, 16, 32, 64, 80, 96, 128, 192, 256 };
    for (int mode = 0; mode < 3; ++mode)
        for (uint32_t n : ns) {
            const uint32_t step = n * 32;          // bytes of B per K step
            if (mode == 0) probe<0><<<1, 128, 96 * 1024>>>(n, step, d_out); else if (mode == 1) probe<1><<<1, 128, 96 * 1024>>>(n, step, d_out); else probe<2><<<1, 128, 96 * 1024>>>(n, step, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
            printf("A in %s N=%3u: %7.1f | %7.1f cycles/MMA  (floor 128*N/256 = %5.1f)  %s%s\n", mode == 0 ? "TMEM" : mode == 1 ? "smem" : "TMEM (converged warp, elect)", n, h[0] / 512.0, h[1] / 512.0,
                   128.0 * n / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e), h[2] ? "" : " TIMEOUT");
        }
    return 0;
}
