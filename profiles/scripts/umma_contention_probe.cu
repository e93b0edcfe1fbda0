// umma_contention_probe.cu — does register -> TMEM traffic (tcgen05.st from the expander warps of local_tc5.cu) or shared-memory
// traffic slow the tensor pipe down?  Warp 1 issues 512 tcgen05.mma.kind::i8 (M 128, N 80, K 32, A in TMEM, converged + elect.sync,
// same accumulator) while warps 4-7 (one per TMEM lane quarter)
//   mode 0: do nothing              mode 1: store 4 KB per warp to OTHER TMEM columns back to back (tcgen05.st.32x32b.x32)
//   mode 2: as 1, with tcgen05.wait::st after every store     mode 3: write shared memory (st.shared.v4) back to back
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I genomicsimulation_b200/csrc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc5_common.cuh"
using namespace gsb::tc5;

__global__ void __launch_bounds__(256, 1) probe(int mode, long long* out, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t s_b[];          // 64 KB of B + 32 KB scratch
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t i = tid; i < 96 * 1024; i += 256) s_b[i] = (uint8_t) (i * 7u);
    if (tid == 0) { mbar_init(&s_bar, 1); s_stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp == 1) {
        const uint32_t idesc = idesc_i8(80), bar = smem_u32(&s_bar), b_base = smem_u32(s_b);
        const uint64_t bdesc0 = b_descriptor(b_base);
        bool ok = true;
        const long long t0 = clock64();
#pragma unroll 1
        for (int trip = 0; trip < 64; ++trip) {
            uint64_t bd = bdesc0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t leader;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
                if (leader) umma_i8_ts(tmem, tmem + 256u + 8u * k, bd, idesc, 1u);
                __syncwarp();
                bd += 160;
            }
        }
        if (lane == 0) {
            tc_commit(bar);
            mbar_wait(bar, 0, ok);
            out[0] = clock64() - t0; out[1] = ok;
            s_stop = 1;
        }
    } else if (warp >= 4 && mode > 0) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = tid * 33u + i;
        const uint32_t taddr = tmem + (((warp & 3) * 32) << 16) + 384u;
        uint32_t n = 0;
        while (!s_stop) {
            if (mode == 3) {
                const uint32_t a = smem_u32(s_b) + 65536u + (tid - 128) * 16u + (n & 7u) * 2048u;
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
            } else {
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                    :: "r"(taddr + 32u * (n & 3u)),
                       "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                       "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                       "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                    : "memory");
                if (mode == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            ++n;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (lane == 0) sink[warp] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d_out; long long h[2]; uint32_t* d_sink; uint32_t hs[8];
    cudaMalloc(&d_out, 16); cudaMalloc(&d_sink, 32);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const char* names[] = { "alone", "4 warps storing to TMEM back to back", "4 warps storing to TMEM, wait::st each", "4 warps writing shared memory" };
    for (int mode = 0; mode < 4; ++mode) {
        cudaMemset(d_sink, 0, 32);
        probe<<<1, 256, 96 * 1024>>>(mode, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost); cudaMemcpy(hs, d_sink, 32, cudaMemcpyDeviceToHost);
        printf("MMA N=80, %-42s: %6.1f cycles per MMA  (side warps did %u + %u + %u + %u ops)  %s%s\n", names[mode], h[0] / 512.0,
               hs[4], hs[5], hs[6], hs[7], e == cudaSuccess ? "" : cudaGetErrorString(e), h[1] ? "" : " TIMEOUT");
    }
    return 0;
}
