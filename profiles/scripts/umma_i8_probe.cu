// umma_i8_probe.cu — what does one tcgen05.mma.kind::i8 (M = 128, K = 32, A in TMEM, B in shared memory) cost
// as a function of N, and does accumulating consecutive MMAs into the SAME accumulator serialise them?
// One CTA, one issuing thread, 512 back-to-back MMAs per configuration, clock64 from first issue to the
// commit's mbarrier completion.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I genomicsimulation_b200/csrc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc5_common.cuh"
using namespace gsb::tc5;

__global__ void __launch_bounds__(128, 1) probe(uint32_t n_cols, uint32_t n_acc, uint32_t a_from_stages, long long* out) {
    extern __shared__ __align__(128) uint8_t s_b[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 32 * 256 * 4; i += 128) s_b[i] = (uint8_t) (i * 7u);
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
    if (tid == 0) {
        const uint32_t idesc = idesc_i8(n_cols), bar = smem_u32(&s_bar), b_base = smem_u32(s_b);
        const int reps = 512;
        bool ok = true;
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            const uint32_t d = tmem + (uint32_t) (i % n_acc) * 64u;                  // accumulators 64 columns apart (n_cols <= 64 when n_acc > 1 ... or overlap, harmless)
            const uint32_t a = tmem + 384u + 8u * (uint32_t) (i % a_from_stages);    // A: 8 columns per K-step
            umma_i8_ts(d, a, b_descriptor(b_base + (uint32_t) (i & 3) * 8192u), idesc, i >= (int) n_acc);
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
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 256 * 4);
    printf("# tcgen05.mma.kind::i8 M=128 K=32, A in TMEM, B no-swizzle K-major in smem; 512 MMAs; cycles per MMA (issue only | until complete)\n");
    const uint32_t ns[] = { 8, 16, 32, 48, 64, 80, 96, 128, 256 };
    for (uint32_t n : ns)
        for (uint32_t acc : { 1u, 2u, 4u }) {
            if (acc > 1 && n > 64) continue;
            for (uint32_t stages : { 1u, 8u }) {
                probe<<<1, 128, 32 * 256 * 4>>>(n, acc, stages, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
                printf("N=%3u accumulators=%u A-steps=%u: %7.1f | %7.1f cycles/MMA %s %s\n", n, acc, stages, h[0] / 512.0, h[1] / 512.0,
                       h[2] ? "" : "(TIMEOUT)", e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    return 0;
}
