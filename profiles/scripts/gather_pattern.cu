// Micro-benchmark: DRAM read bandwidth of the two ways a block of 4 warps can walk packed genotype
// rows for the GEBV mma (loads only, same lane mapping and prefetch depth as gebv_counts_kernel).
//   mode 0: every warp owns 16 individuals and walks their 32 strand rows 64 bytes at a time
//   mode 1: the block owns 16 individuals; warp w takes chunks w, w + 4, ... so the block requests
//           256 contiguous bytes of every row at a time
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_pattern gather_pattern.cu && ./gather_pattern
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int kMode>
__global__ void __launch_bounds__(128, 7) walk(const uint4* __restrict__ store, uint32_t row_u4, uint32_t plane_u4, uint32_t n,
                                               uint32_t n_segments, uint32_t seg_chunks, uint32_t chunks, uint32_t* out) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const uint32_t seg = blockIdx.x % n_segments, gb = blockIdx.x / n_segments;
    const uint32_t group = kMode == 0 ? gb * 4 + warp : gb;
    const uint32_t ch_lo = seg * seg_chunks, ch_hi = min(ch_lo + seg_chunks, chunks);
    const uint32_t ia = min(group * 16 + g, n - 1), ib = min(group * 16 + g + 8, n - 1);
    const uint4* pa = store + (size_t) ia * row_u4 + c;
    const uint4* pb = store + (size_t) ib * row_u4 + c;
    const uint32_t first = kMode == 0 ? ch_lo : ch_lo + warp, step = kMode == 0 ? 1 : 4;
    uint32_t acc = 0;
    if (first >= ch_hi) return;
    uint4 x0 = ldg_stream(pa + first * 4), x1 = ldg_stream(pa + plane_u4 + first * 4), x2 = ldg_stream(pb + first * 4), x3 = ldg_stream(pb + plane_u4 + first * 4);
    for (uint32_t ch = first; ch < ch_hi; ch += step) {
        const uint32_t nx = min(ch + step, ch_hi - 1) * 4;
        const uint4 y0 = ldg_stream(pa + nx), y1 = ldg_stream(pa + plane_u4 + nx), y2 = ldg_stream(pb + nx), y3 = ldg_stream(pb + plane_u4 + nx);
        acc ^= x0.x ^ x0.y ^ x0.z ^ x0.w ^ x1.x ^ x1.y ^ x1.z ^ x1.w ^ x2.x ^ x2.y ^ x2.z ^ x2.w ^ x3.x ^ x3.y ^ x3.z ^ x3.w;
        // stand-in for the chunk's arithmetic: ~250 dependent-free cycles
        for (int k = 0; k < 60; ++k) acc = acc * 1664525u + 1013904223u;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    const uint32_t n = 100000, plane_words = 1568, row_u4 = plane_words * 2 / 4, plane_u4 = plane_words / 4, chunks = plane_words / 16;
    uint4* store; uint32_t* out;
    cudaMalloc(&store, (size_t) n * row_u4 * 16); cudaMalloc(&out, 4);
    cudaMemset(store, 1, (size_t) n * row_u4 * 16);
    const double bytes = (double) n * row_u4 * 16;
    for (int mode = 0; mode < 2; ++mode) {
        for (uint32_t n_segments : {1u, 2u, 4u, 7u}) {
            const uint32_t seg_chunks = (chunks + n_segments - 1) / n_segments;
            const uint32_t groups = (n + 15) / 16, gb = mode == 0 ? (groups + 3) / 4 : groups;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) walk<0><<<gb * n_segments, 128>>>(store, row_u4, plane_u4, n, n_segments, seg_chunks, chunks, out);
                else walk<1><<<gb * n_segments, 128>>>(store, row_u4, plane_u4, n, n_segments, seg_chunks, chunks, out);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (rep && ms < best) best = ms;
            }
            printf("mode %d  segments %u: %.4f ms  %.0f GB/s\n", mode, n_segments, best, bytes / best / 1e6);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
