// tmem_bw_probe.cu — what do the register <-> TMEM moves of local_tc5.cu cost?  Four warps (one per
// TMEM lane quarter), each issuing 256 tcgen05.st.32x32b.x32 (4 KB per warp instruction) or
// tcgen05.ld.32x32b.x8 / .x32 back to back; cycles per instruction and bytes per cycle per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I genomicsimulation_b200/csrc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc5_common.cuh"
using namespace gsb::tc5;

__global__ void __launch_bounds__(128, 1) probe(int mode, int n_warps, long long* out, uint32_t* sink) {
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem + ((warp * 32) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if ((int) warp < n_warps) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = tid * 33u + i;
        t0 = clock64();
        if (mode == 0) {
#pragma unroll 1
            for (int it = 0; it < 64; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                        :: "r"(tmem + 32u * k),
                           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                        : "memory");
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < 64; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t d[8];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
                                 : "r"(tmem + 8u * k) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    acc += d[0] ^ d[7];
                }
            }
        }
        t1 = clock64();
    }
    if ((tid & 31) == 0) { out[warp * 2] = t0; out[warp * 2 + 1] = t1; }
    sink[tid] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_out; long long h[8]; uint32_t* d_sink;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 512);
    for (int mode = 0; mode < 2; ++mode)
        for (int nw = 1; nw <= 4; nw *= 2) {
            probe<<<1, 128>>>(mode, nw, d_out, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
            long long lo = h[0], hi = h[1];
            for (int w = 0; w < nw; ++w) { if (h[2 * w] < lo) lo = h[2 * w]; if (h[2 * w + 1] > hi) hi = h[2 * w + 1]; }
            const double cyc = (double) (hi - lo), per = cyc / 256.0;
            const double bytes = (mode == 0 ? 4096.0 : 1024.0) * 256.0 * nw;
            printf("%s, %d warp(s): %7.1f cycles per warp instruction, %6.1f B/cycle/SM  %s\n",
                   mode == 0 ? "tcgen05.st.32x32b.x32 (4 KB)" : "tcgen05.ld.32x32b.x8 (1 KB) + wait::ld", nw, per, bytes / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
