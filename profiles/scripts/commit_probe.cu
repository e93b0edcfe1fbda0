// commit_probe.cu — what does tcgen05.commit cost the issuing thread?  One thread, 256 trips of
//   mode 0: commit -> mbarrier wait (round trip)
//   mode 1: 8 MMAs (N = 80) + commit, barrier waited for 4 trips later (the pipelined use of local_tc5.cu)
//   mode 2: commit only, to 8 rotating barriers nobody waits for
//   mode 3: as 1, the thread running inside `if (lane == 0)` of a full warp (divergent) instead of a 1-thread role
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I genomicsimulation_b200/csrc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc5_common.cuh"
using namespace gsb::tc5;

__global__ void __launch_bounds__(128, 1) probe(int mode, long long* out) {
    extern __shared__ __align__(128) uint8_t s_b[];
    __shared__ __align__(8) uint64_t s_bar[8];
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 64 * 1024; i += 128) s_b[i] = (uint8_t) (i * 7u);
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&s_bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp == 1 && (tid & 31) == 0) {
        const uint32_t idesc = idesc_i8(80), bar = smem_u32(s_bar), b_base = smem_u32(s_b);
        const uint64_t bdesc0 = b_descriptor(b_base);
        bool ok = true;
        const long long t0 = clock64();
#pragma unroll 1
        for (int trip = 0; trip < 256; ++trip) {
            const uint32_t k = trip & 7;
            if (mode == 0) {
                tc_commit(bar + 8 * k);
                mbar_wait(bar + 8 * k, (trip >> 3) & 1, ok);
            } else if (mode == 1 || mode == 3) {
                if (trip >= 4) mbar_wait(bar + 8 * ((trip - 4) & 7), ((trip - 4) >> 3) & 1, ok);
                uint64_t bd = bdesc0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { umma_i8_ts(tmem, tmem + 256u + 8u * j, bd, idesc, 1u); bd += 160; }
                tc_commit(bar + 8 * k);
            } else {
                tc_commit(bar + 8 * k);
            }
        }
        const long long t1 = clock64();
        out[0] = t1 - t0; out[1] = ok;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_out; long long h[2];
    cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[] = { "commit + wait round trip", "8 MMA (N=80) + commit, waited 4 trips later", "commit only" };
    for (int mode = 0; mode < 3; ++mode) {
        probe<<<1, 128, 64 * 1024>>>(mode, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%-46s: %8.1f cycles per trip  %s%s\n", names[mode], h[0] / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e), h[1] ? "" : " TIMEOUT");
    }
    return 0;
}
