// Microbenchmark behind DESIGN.md §3.2: do the bit->byte expansion (LOP3, alu pipe) and the legacy
// integer MMA (IMMA.16832.U8.S8) overlap on an sm_100a scheduler, or do they share issue bandwidth?
// Each variant runs `iters` iterations of the GEBV inner step shape (8 LOP3 + 2 IMMA per marker
// word) with 1..8 warps per scheduler and prints cycles per iteration per scheduler.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_overlap pipe_overlap.cu && ./pipe_overlap
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_u8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// mode 0: LOP3 only (8 per iteration, register mask)      mode 1: IMMA only (2 per iteration)
// mode 2: both, IMMA consumes the LOP3 results            mode 3: both, independent
// mode 4: LOP3 only, immediate mask                        mode 5: 4 IMMA chains only (latency check)
template <int kMode>
__global__ void probe(uint32_t* out, uint32_t iters, uint32_t seed) {
    const uint32_t lane = threadIdx.x & 31, c = lane & 3;
    const uint32_t k_lo = 0x01010101u << c, k_hi = k_lo << 4;
    uint32_t w0 = seed * 2654435761u + threadIdx.x, w1 = w0 ^ 0x9e3779b9u, w2 = w0 * 3u, w3 = w1 * 5u;
    int acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0}, acc2[4] = {0, 0, 0, 0}, acc3[4] = {0, 0, 0, 0};
    uint32_t sink = 0;
    const uint32_t b0 = seed | 0x01020304u, b1 = seed | 0x04030201u;
    const long long t0 = clock64();
#pragma unroll 4
    for (uint32_t i = 0; i < iters; ++i) {
        // the "loads": cheap per-iteration changes so that nothing hoists
        w0 = w0 * 0x01000193u + 7u; w1 = w1 * 0x01000193u + 7u; w2 = w2 * 0x01000193u + 7u; w3 = w3 * 0x01000193u + 7u;   // IMAD: fma pipe
        if (kMode == 0) {
            sink ^= (w0 & k_lo) ^ (w1 & k_lo) ^ (w0 & k_hi) ^ (w1 & k_hi) ^ (w2 & k_lo) ^ (w3 & k_lo) ^ (w2 & k_hi) ^ (w3 & k_hi);
        } else if (kMode == 4) {
            sink ^= (w0 & 0x01010101u) ^ (w1 & 0x02020202u) ^ (w0 & 0x10101010u) ^ (w1 & 0x20202020u)
                  ^ (w2 & 0x04040404u) ^ (w3 & 0x08080808u) ^ (w2 & 0x40404040u) ^ (w3 & 0x80808080u);
        } else if (kMode == 1) {
            mma_u8s8(acc0, w0, w1, w0, w1, b0, b1);
            mma_u8s8(acc1, w2, w3, w2, w3, b0, b1);
        } else if (kMode == 2) {
            mma_u8s8(acc0, w0 & k_lo, w1 & k_lo, w0 & k_hi, w1 & k_hi, b0, b1);
            mma_u8s8(acc1, w2 & k_lo, w3 & k_lo, w2 & k_hi, w3 & k_hi, b0, b1);
        } else if (kMode == 3) {
            sink ^= (w0 & k_lo) ^ (w1 & k_lo) ^ (w0 & k_hi) ^ (w1 & k_hi) ^ (w2 & k_lo) ^ (w3 & k_lo) ^ (w2 & k_hi) ^ (w3 & k_hi);
            mma_u8s8(acc0, b0, b1, b0, b1, b0, b1);
            mma_u8s8(acc1, b1, b0, b1, b0, b0, b1);
        } else if (kMode == 5) {
            mma_u8s8(acc0, w0, w1, w0, w1, b0, b1);
            mma_u8s8(acc1, w2, w3, w2, w3, b0, b1);
            mma_u8s8(acc2, w1, w0, w1, w0, b0, b1);
            mma_u8s8(acc3, w3, w2, w3, w2, b0, b1);
        }
    }
    const long long t1 = clock64();
    for (int q = 0; q < 4; ++q) sink ^= acc0[q] ^ acc1[q] ^ acc2[q] ^ acc3[q];
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 2] = sink;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 2 + 1] = (uint32_t) (t1 - t0);
}

template <int kMode>
static void run(const char* name, uint32_t* d_out, uint32_t* h_out) {
    const uint32_t iters = 4096;
    printf("%-44s", name);
    for (int wps = 1; wps <= 8; wps *= 2) {          // warps per scheduler
        const int threads = 4 * wps * 32;
        probe<kMode><<<148, threads>>>(d_out, iters, 12345u);
        probe<kMode><<<148, threads>>>(d_out, iters, 12345u);
        cudaDeviceSynchronize();
        cudaMemcpy(h_out, d_out, 148 * threads * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        double cyc = 0; for (int t = 0; t < 148 * threads; ++t) cyc += h_out[2 * t + 1];
        cyc /= 148.0 * threads;
        // every warp ran `iters` iterations; the scheduler served wps warps in `cyc` cycles
        printf("  wps=%d: %6.2f", wps, cyc / iters / wps);
    }
    printf("   cycles / iteration / scheduler\n");
}

int main() {
    uint32_t* d_out; cudaMalloc(&d_out, 148 * 1024 * 2 * sizeof(uint32_t));
    uint32_t* h_out = (uint32_t*) malloc(148 * 1024 * 2 * sizeof(uint32_t));
    run<0>("8 LOP3 (register mask)", d_out, h_out);
    run<4>("8 LOP3 (immediate mask)", d_out, h_out);
    run<1>("2 IMMA.16832.U8.S8", d_out, h_out);
    run<5>("4 IMMA.16832.U8.S8", d_out, h_out);
    run<2>("8 LOP3 -> 2 IMMA (dependent, the kernel)", d_out, h_out);
    run<3>("8 LOP3 + 2 IMMA (independent)", d_out, h_out);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
