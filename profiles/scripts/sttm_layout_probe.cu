// Which (thread, register) of a tcgen05.st.16x256b lands in which (TMEM lane, column)?  Writes
// tagged words with the 16x256b shape at lane offsets 0 and 16 of each warp's quadrant and reads
// them back with the 32x32b shape (thread = lane, register = column).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sttm_probe sttm_layout_probe.cu && ./sttm_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t) __cvta_generic_to_shared(&s_tmem)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t r[8];
    for (int h = 0; h < 2; ++h) {
        for (int q = 0; q < 8; ++q) r[q] = (h << 16) | (lane << 8) | q;      // tag: half, thread, register
        asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(tmem + ((warp * 32 + h * 16) << 16)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t d[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                   "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                 : "r"(tmem + ((warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[tid * 16 + j] = d[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32u) : "memory");
}

int main() {
    uint32_t* d; cudaMalloc(&d, 128 * 16 * 4);
    cudaMemset(d, 0xff, 128 * 16 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t h[128 * 16];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("status: %s\nTMEM lane : column -> half.thread.register that wrote it\n", cudaGetErrorString(e));
    for (int l = 0; l < 128; ++l) {
        if (l >= 34 && l < 126) continue;
        printf("lane %3d:", l);
        for (int j = 0; j < 16; ++j) printf(" %x.%02u.%u", h[l * 16 + j] >> 16, (h[l * 16 + j] >> 8) & 255, h[l * 16 + j] & 255);
        printf("\n");
    }
    return e != cudaSuccess;
}
