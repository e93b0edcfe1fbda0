// ipc_probe.cu — does CUDA IPC peer memory work between two processes on this box, and what do
// P2P pushes over NVLink cost?  Two processes (fork before any CUDA call), one GPU each
// (both on GPU 0 when only one is visible), exchange cudaIpcMemHandles over pipes, then:
//   1. each pushes 64 MB into the peer's slab with plain 128-bit stores, fences, raises a flag
//      in the peer's memory; the peer spins on the flag (bounded) and verifies the data;
//   2. ping-pong of flags: one-way latency of a release store -> acquire spin;
//   3. push bandwidth for 1 MB / 16 MB / 128 MB.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ipc_probe ipc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("[rank %d] %s -> %s\n", g_rank, #x, cudaGetErrorString(e_)); fflush(stdout); _exit(3); } } while (0)
static int g_rank = 0;

__global__ void push_kernel(uint4* peer, const uint4* mine, size_t n16, unsigned* counter, volatile unsigned* peer_flag, unsigned epoch) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x) peer[i] = mine[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            __threadfence_system();
            *peer_flag = epoch;
            *counter = 0;
        }
    }
}

__global__ void wait_kernel(volatile unsigned* flag, unsigned epoch, int* timeout) {
    const long long t0 = clock64();
    while (*flag < epoch) {
        if (clock64() - t0 > 4000000000ll) { *timeout = 1; return; }
    }
    __threadfence_system();
}

__global__ void fill_kernel(uint4* p, size_t n16, unsigned seed) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x)
        p[i] = make_uint4((unsigned) i ^ seed, seed, (unsigned) (i >> 3), 7u);
}

__global__ void check_kernel(const uint4* p, size_t n16, unsigned seed, unsigned long long* bad) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x) {
        const uint4 v = p[i];
        if (v.x != ((unsigned) i ^ seed) || v.y != seed || v.z != (unsigned) (i >> 3) || v.w != 7u) atomicAdd(bad, 1ull);
    }
}

struct Handle { cudaIpcMemHandle_t h; int pid; int device; };

int main() {
    int to_child[2], to_parent[2];
    if (pipe(to_child) || pipe(to_parent)) return 2;
    const pid_t pid = fork();
    g_rank = pid == 0 ? 1 : 0;
    const int rd = g_rank ? to_child[0] : to_parent[0], wr = g_rank ? to_parent[1] : to_child[1];
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const int dev = g_rank % ndev;
    CK(cudaSetDevice(dev));
    const size_t slab = (size_t) 160 << 20, data_off = 4096;
    char* mine = nullptr;
    CK(cudaMalloc(&mine, slab));
    CK(cudaMemset(mine, 0, slab));
    char* src = nullptr;
    CK(cudaMalloc(&src, slab));
    Handle me, other;
    memset(&me, 0, sizeof me);
    CK(cudaIpcGetMemHandle(&me.h, mine));
    me.pid = getpid(); me.device = dev;
    if (write(wr, &me, sizeof me) != (ssize_t) sizeof me || read(rd, &other, sizeof other) != (ssize_t) sizeof other) return 2;
    char* peer = nullptr;
    CK(cudaIpcOpenMemHandle((void**) &peer, other.h, cudaIpcMemLazyEnablePeerAccess));
    printf("[rank %d] device %d of %d, peer device %d: IPC handle opened\n", g_rank, dev, ndev, other.device); fflush(stdout);
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    unsigned* counter; int* timeout; unsigned long long* bad;
    CK(cudaMalloc(&counter, 4)); CK(cudaMemset(counter, 0, 4));
    CK(cudaMalloc(&timeout, 4)); CK(cudaMemset(timeout, 0, 4));
    CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    CK(cudaDeviceSynchronize());
    char token = 1;
    if (write(wr, &token, 1) != 1 || read(rd, &token, 1) != 1) return 2;     // both slabs are zeroed and mapped

    // 1. push 64 MB + flag, peer verifies
    const size_t n16 = ((size_t) 64 << 20) / 16;
    unsigned epoch = 1;
    fill_kernel<<<592, 256, 0, s>>>((uint4*) src, n16, 0x1000u + g_rank);
    push_kernel<<<592, 256, 0, s>>>((uint4*) (peer + data_off), (const uint4*) src, n16, counter, (volatile unsigned*) peer, epoch);
    wait_kernel<<<1, 1, 0, s>>>((volatile unsigned*) mine, epoch, timeout);
    check_kernel<<<592, 256, 0, s>>>((const uint4*) (mine + data_off), n16, 0x1000u + (1 - g_rank), bad);
    CK(cudaStreamSynchronize(s));
    int h_to = 0; unsigned long long h_bad = 0;
    CK(cudaMemcpy(&h_to, timeout, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&h_bad, bad, 8, cudaMemcpyDeviceToHost));
    printf("[rank %d] push+flag: timeout=%d mismatches=%llu\n", g_rank, h_to, h_bad); fflush(stdout);

    // 2. flag ping-pong latency: rank 0 raises, rank 1 answers, 200 round trips in ONE pair of streams
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (write(wr, &token, 1) != 1 || read(rd, &token, 1) != 1) return 2;
    const int trips = 200;
    CK(cudaEventRecord(e0, s));
    for (int t = 0; t < trips; ++t) {
        ++epoch;
        if (g_rank == 0) {
            push_kernel<<<1, 32, 0, s>>>((uint4*) (peer + data_off), (const uint4*) src, 0, counter, (volatile unsigned*) peer, epoch);
            wait_kernel<<<1, 1, 0, s>>>((volatile unsigned*) mine, epoch, timeout);
        } else {
            wait_kernel<<<1, 1, 0, s>>>((volatile unsigned*) mine, epoch, timeout);
            push_kernel<<<1, 32, 0, s>>>((uint4*) (peer + data_off), (const uint4*) src, 0, counter, (volatile unsigned*) peer, epoch);
        }
    }
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(&h_to, timeout, 4, cudaMemcpyDeviceToHost));
    printf("[rank %d] %d flag round trips (2 launches each side per trip): %.3f ms -> %.2f us per round trip, timeout=%d\n",
           g_rank, trips, ms, ms * 1e3 / trips, h_to); fflush(stdout);

    // 3. push bandwidth
    const size_t sizes[3] = { (size_t) 1 << 20, (size_t) 16 << 20, (size_t) 128 << 20 };
    for (int k = 0; k < 3; ++k) {
        if (write(wr, &token, 1) != 1 || read(rd, &token, 1) != 1) return 2;
        const size_t m16 = sizes[k] / 16;
        for (int rep = 0; rep < 3; ++rep) {
            ++epoch;
            CK(cudaEventRecord(e0, s));
            push_kernel<<<592, 256, 0, s>>>((uint4*) (peer + data_off), (const uint4*) src, m16, counter, (volatile unsigned*) peer, epoch);
            wait_kernel<<<1, 1, 0, s>>>((volatile unsigned*) mine, epoch, timeout);
            CK(cudaEventRecord(e1, s));
            CK(cudaStreamSynchronize(s));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 2) printf("[rank %d] push %zu MB both ways + flag: %.3f ms -> %.1f GB/s per direction\n", g_rank, sizes[k] >> 20, ms, sizes[k] / ms / 1e6);
        }
        fflush(stdout);
    }
    CK(cudaIpcCloseMemHandle(peer));
    if (g_rank == 0) { int st = 0; waitpid(pid, &st, 0); printf("child exit %d\n", WEXITSTATUS(st)); }
    return 0;
}
