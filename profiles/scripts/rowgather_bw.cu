// Micro-benchmark: read bandwidth of the GEBV access pattern.  N rows of ROW bytes (the packed
// store: 12 544 B per individual); a warp owns 32 strand rows (16 individuals x 2 strands, each
// 6 272 B) and walks them in lock step, R contiguous bytes per row per step — the pattern an mma
// tile needs.  Prints GB/s for R = 16 .. 1024 and for a plain row-major stream.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rowgather_bw rowgather_bw.cu && ./rowgather_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void gather(const uint4* __restrict__ base, uint32_t n_groups, uint32_t plane_u4, uint32_t r_u4, uint32_t* out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_groups) return;
    // 32 strand rows of this warp: row r starts at (warp * 32 + r) * plane_u4
    const uint32_t lanes_per_row = r_u4 < 32 ? r_u4 : 32;            // lanes cooperating on one row's R bytes
    const uint32_t rows_per_instr = 32 / lanes_per_row;
    uint32_t acc = 0;
    for (uint32_t off = 0; off < plane_u4; off += r_u4) {
        for (uint32_t r0 = 0; r0 < 32; r0 += rows_per_instr) {
            const uint32_t row = r0 + lane / lanes_per_row;
            for (uint32_t p = lane % lanes_per_row; p < r_u4; p += lanes_per_row) {
                if (off + p < plane_u4) {
                    const uint4 v = __ldg(base + (size_t) (warp * 32 + row) * plane_u4 + off + p);
                    acc ^= v.x ^ v.y ^ v.z ^ v.w;
                }
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void stream(const uint4* __restrict__ base, size_t n_u4, uint32_t* out) {
    uint32_t acc = 0;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_u4; i += (size_t) gridDim.x * blockDim.x) {
        const uint4 v = __ldg(base + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    const uint32_t n_ind = 100000, plane_bytes = 6272, plane_u4 = plane_bytes / 16;
    const size_t bytes = (size_t) n_ind * 2 * plane_bytes;
    uint4* d; uint32_t* out;
    cudaMalloc(&d, bytes); cudaMalloc(&out, 4); cudaMemset(d, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const uint32_t n_groups = n_ind * 2 / 32;
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a); stream<<<148 * 16, 256>>>(d, bytes / 16, out); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    printf("row-major stream: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    for (uint32_t r_bytes = 16; r_bytes <= 2048; r_bytes *= 2) {
        for (int threads = 128; threads <= 256; threads *= 2) {
            const uint32_t blocks = (n_groups * 32 + threads - 1) / threads;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a); gather<<<blocks, threads>>>(d, n_groups, plane_u4, r_bytes / 16, out); cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
            }
            printf("R = %4u B per row per step, %3d threads/block: %.3f ms  %.0f GB/s\n", r_bytes, threads, ms, bytes / ms / 1e6);
        }
    }
    return 0;
}
