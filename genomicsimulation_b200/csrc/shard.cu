// shard.cu — one generation of recurrent selection sharded over the GPUs of a box.
//
// One process (or context) per GPU holds a contiguous index range of the generation.  Per
// generation, all on the context stream and with NO host synchronisation:
//
//   select:  GEBV of the local individuals (evaluate.cu)            gsc_calculate_bvs   core.c:9536
//            -> pushed into every peer's copy of the generation's GEBV vector (P2P stores over
//               NVLink + a release flag per rank)                    (the "all-gather")
//            -> replicated radix selection of the global top_n (select.cu): every rank computes
//               the same ascending list of selected global indexes   gsc_split_by_bv     core.c:9463-9517
//            -> each rank copies the selected rows it owns out of its packed store straight into
//               every peer's parent-pool buffer, at their rank in the selection (the "pool
//               broadcast"), then raises its pool flag
//   cross:   waits for every rank's pool flag, draws random pairs over the pool on the device
//            (pick rule of core.c:8556-8581, Philox position = GLOBAL offspring number) and
//            splices this rank's range of the next generation (meiosis.cu)  gsc_make_random_crosses core.c:8483
//
// The exchange buffers of a rank are one cudaMalloc slab, opened by its peers through CUDA IPC
// (or used directly when two shards live in one process).  Why not NCCL here: the number of pool
// rows a rank contributes is known only on the device (it depends on the selection), so an
// all-gather-v would need a host round trip per generation; the push kernels need none, fuse the
// row gather with the transfer, and are latency-bound at ~7 us per flag hop instead of ~60 us per
// small NCCL collective (profiles/r2_ipc_probe_2gpu.txt: 642 GB/s per direction for 128 MB).
//
// Buffer reuse is safe without extra barriers: a peer can only push generation g+1's GEBVs after its
// cross(g), which waited for MY pool flag of g, which I raise after my selection of g has read the
// GEBV vector; and it can only push generation g+1's pool rows after its selection of g+1, which
// waited for MY GEBV flag of g+1, which I raise after my cross(g) has read the pool.
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cstring>
#include <unistd.h>

using namespace gsb;

namespace {

constexpr uint32_t kMaxWorld = 16;
constexpr uint32_t kHandleMagic = 0x67736232u;
constexpr uint32_t kTimingRing = 256, kTimingEvents = 7;
constexpr long long kWaitTimeoutCycles = 40000000000ll;      // ~20 s at 1.9 GHz: a peer that never shows up

struct Peers { char* base[kMaxWorld]; uint32_t row_u4; };      // row_u4: uint4 per packed row

struct ShardHandle {                 // what gsb200_shard_handle exports (GSB200_SHARD_HANDLE_BYTES)
    uint32_t magic, rank, world, device;
    int32_t pid;
    uint32_t max_pool;
    uint64_t slab_bytes, row_bytes, max_total;
    void* raw;                       // the slab's address in its owner's process
    cudaIpcMemHandle_t ipc;
    uint32_t pci;                    // which physical GPU: domain << 16 | bus << 8 | device
};
static_assert(sizeof(ShardHandle) <= GSB200_SHARD_HANDLE_BYTES, "shard handle too large");

// Layout of a slab.  flags[0][r]: epoch of rank r's GEBVs present here; flags[1][r]: of its pool rows.
constexpr size_t kOffFlags = 0, kOffBv = 512;

__device__ __forceinline__ void raise_flags(const Peers& P, uint32_t world, uint32_t rank, uint32_t which, uint32_t epoch) {
    for (uint32_t r = 0; r < world; ++r)
        reinterpret_cast<volatile uint32_t*>(P.base[r] + kOffFlags)[which * kMaxWorld + rank] = epoch;
}

// my GEBVs (and ids) -> every peer's vector; the last block to finish raises my GEBV flag everywhere
__global__ void __launch_bounds__(256) shard_push_values_kernel(const __grid_constant__ Peers P, uint32_t world, uint32_t rank, size_t off_ids,
                                                                uint32_t lo, uint32_t n_local, int with_ids,
                                                                uint32_t* counter, uint32_t epoch) {
    const double* bv = reinterpret_cast<const double*>(P.base[rank] + kOffBv) + lo;
    const uint32_t* ids = reinterpret_cast<const uint32_t*>(P.base[rank] + off_ids) + lo;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += gridDim.x * blockDim.x) {
        const double v = bv[i];
        const uint32_t id = with_ids ? ids[i] : 0u;
        for (uint32_t r = 0; r < world; ++r) {
            if (r == rank) continue;
            (reinterpret_cast<double*>(P.base[r] + kOffBv) + lo)[i] = v;
            if (with_ids) (reinterpret_cast<uint32_t*>(P.base[r] + off_ids) + lo)[i] = id;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(counter, 1u) == gridDim.x - 1) {
        __threadfence_system();
        raise_flags(P, world, rank, 0, epoch);
        *counter = 0;
    }
}

// one lane per rank spins on that rank's flag in MY slab
__global__ void shard_wait_kernel(const volatile uint32_t* flags, uint32_t world, uint32_t epoch, int* err) {
    if (threadIdx.x < world) {
        const long long t0 = clock64();
        while (flags[threadIdx.x] < epoch) {
            if (clock64() - t0 > kWaitTimeoutCycles) { atomicExch(err, 1); break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// the selected rows I own -> every rank's pool buffer (mine included), at their rank in the selection;
// ids of the whole selection (local reads); the last block raises my pool flag everywhere
__global__ void __launch_bounds__(256) shard_push_pool_kernel(const __grid_constant__ Peers P, uint32_t world, uint32_t rank, size_t off_ids, size_t off_pool,
                                                              const uint32_t* __restrict__ store, uint64_t row_words,
                                                              const uint32_t* __restrict__ local_rows, uint32_t lo,
                                                              const SelectState* __restrict__ state, const uint32_t* __restrict__ sel,
                                                              uint32_t top_n, uint32_t* __restrict__ sel_ids, int with_ids,
                                                              uint32_t* counter, uint32_t epoch) {
    const uint32_t own_lo = state->own_lo, own_cnt = state->own_hi - own_lo;
    const uint32_t pieces = (uint32_t) (row_words / 4);
    const uint64_t total = (uint64_t) own_cnt * pieces;
    for (uint64_t flat = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; flat < total; flat += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t p = own_lo + (uint32_t) (flat / pieces), w = (uint32_t) (flat % pieces);
        const uint32_t row = local_rows[sel[p] - lo];
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(store + (uint64_t) row * row_words) + w);
        for (uint32_t r = 0; r < world; ++r)
            reinterpret_cast<uint4*>(P.base[r] + off_pool)[(uint64_t) p * pieces + w] = v;
    }
    const uint32_t* all_ids = reinterpret_cast<const uint32_t*>(P.base[rank] + off_ids);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < top_n; p += gridDim.x * blockDim.x)
        sel_ids[p] = with_ids ? all_ids[sel[p]] : 0u;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(counter, 1u) == gridDim.x - 1) {
        __threadfence_system();
        raise_flags(P, world, rank, 1, epoch);
        *counter = 0;
    }
}

// The same in two parts, for ranks on different GPUs (see shard_select_impl): the rows I own into MY pool buffer and my flag
// raised here only ...
__global__ void __launch_bounds__(256) shard_push_pool_local_kernel(const __grid_constant__ Peers P, uint32_t rank, size_t off_ids, size_t off_pool,
                                                                    const uint32_t* __restrict__ store, uint64_t row_words,
                                                                    const uint32_t* __restrict__ local_rows, uint32_t lo,
                                                                    const SelectState* __restrict__ state, const uint32_t* __restrict__ sel,
                                                                    uint32_t top_n, uint32_t* __restrict__ sel_ids, int with_ids,
                                                                    uint32_t* counter, uint32_t epoch) {
    const uint32_t own_lo = state->own_lo, own_cnt = state->own_hi - own_lo;
    const uint32_t pieces = (uint32_t) (row_words / 4);
    const uint64_t total = (uint64_t) own_cnt * pieces;
    uint4* mine = reinterpret_cast<uint4*>(P.base[rank] + off_pool);
    for (uint64_t flat = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; flat < total; flat += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t p = own_lo + (uint32_t) (flat / pieces), w = (uint32_t) (flat % pieces);
        const uint32_t row = local_rows[sel[p] - lo];
        mine[(uint64_t) p * pieces + w] = __ldg(reinterpret_cast<const uint4*>(store + (uint64_t) row * row_words) + w);
    }
    const uint32_t* all_ids = reinterpret_cast<const uint32_t*>(P.base[rank] + off_ids);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < top_n; p += gridDim.x * blockDim.x)
        sel_ids[p] = with_ids ? all_ids[sel[p]] : 0u;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(counter, 1u) == gridDim.x - 1) {
        __threadfence();
        reinterpret_cast<volatile uint32_t*>(P.base[rank] + kOffFlags)[1 * kMaxWorld + rank] = epoch;
        *counter = 0;
    }
}

// ... and, on a second stream beside the splice, to one peer after the other: rank + 1 first, so that what a rank receives
// arrives in the order its splice asks for it (gamete_keys_kernel).  The copies are bulk-asynchronous (TMA): ONE thread per
// SM loads 16 KB chunks of my rows into shared memory and stores them to the peer, two chunks in flight — a kernel of
// ordinary loads and stores beside the splice took its issue slots and both crawled (profiles/r2_split_push.txt).  One launch,
// one small block per SM, all resident at once: once started it needs no further SM slots, whatever spins beside it.
constexpr uint32_t kPushChunk = 16384, kPushStages = 2;
__global__ void __launch_bounds__(32) shard_push_pool_remote_kernel(const __grid_constant__ Peers P, uint32_t world, uint32_t rank, size_t off_pool,
                                                                    const uint32_t* __restrict__ bounds /* shard_owner_bounds_kernel */,
                                                                    uint32_t* counters /* [kMaxWorld], zero */, uint32_t epoch) {
    __shared__ __align__(128) unsigned char s_buf[kPushStages][kPushChunk];
    __shared__ __align__(8) unsigned long long s_full[kPushStages];
    const uint32_t own_lo = bounds[rank], own_cnt = bounds[rank + 1] - own_lo;
    const uint64_t row_bytes = (uint64_t) P.row_u4 * 16, total = (uint64_t) own_cnt * row_bytes;       // a multiple of 16
    const char* mine = P.base[rank] + off_pool + (uint64_t) own_lo * row_bytes;
    const uint64_t n_chunks = (total + kPushChunk - 1) / kPushChunk;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < kPushStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t) __cvta_generic_to_shared(&s_full[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t uses[kPushStages] = { 0, 0 };
        for (uint32_t k = 1; k < world; ++k) {
            const uint32_t dest = (rank + k) % world;
            char* theirs = P.base[dest] + off_pool + (uint64_t) own_lo * row_bytes;
            uint32_t st = 0;
            for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, st ^= 1u) {
                const uint64_t at = c * kPushChunk;
                const uint32_t bytes = (uint32_t) min((uint64_t) kPushChunk, total - at);
                const uint32_t buf = (uint32_t) __cvta_generic_to_shared(s_buf[st]), bar = (uint32_t) __cvta_generic_to_shared(&s_full[st]);
                // the store that read this stage two chunks ago has finished READING it
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(buf), "l"(mine + at), "r"(bytes), "r"(bar) : "memory");
                uint32_t done = 0;
                const uint32_t parity = uses[st] & 1u;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                ++uses[st];
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(theirs + at), "r"(buf), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // this block's part of the peer's rows is written
            asm volatile("fence.proxy.async;" ::: "memory");                // ... by the async proxy: ordered before the flag store below
            __threadfence_system();
            if (atomicAdd(&counters[dest], 1u) == gridDim.x - 1) {
                __threadfence_system();
                reinterpret_cast<volatile uint32_t*>(P.base[dest] + kOffFlags)[1 * kMaxWorld + rank] = epoch;
                counters[dest] = 0;
            }
        }
    }
}

// bounds[r] = first slot of the selection (ascending global indexes) owned by rank r; bounds[world] = top_n
__global__ void shard_owner_bounds_kernel(const uint32_t* __restrict__ sel, uint32_t top_n, uint64_t n_total, uint32_t world, uint32_t* __restrict__ bounds) {
    const uint32_t r = threadIdx.x;
    if (r > world) return;
    if (r == world) { bounds[r] = top_n; return; }
    const uint64_t base = n_total / world, extra = n_total % world;
    const uint64_t lo = r * base + (r < extra ? r : extra);          // gsb200_shard_range
    uint32_t a = 0, b = top_n;
    while (a < b) { const uint32_t m = (a + b) / 2; if (sel[m] < lo) a = m + 1; else b = m; }
    bounds[r] = a;
}

__global__ void shard_iota_kernel(uint32_t* out, uint32_t first, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + i;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct gsb200_shard;
static void mark(gsb200_shard* s, uint32_t epoch, uint32_t which);

struct gsb200_shard {
    gsb200_ctx* ctx = nullptr;
    uint32_t rank = 0, world = 1;
    uint64_t max_total = 0;
    uint32_t max_pool = 0;
    uint64_t row_words = 0;
    char* slab = nullptr;
    size_t slab_bytes = 0, off_ids = 0, off_pool = 0;
    Peers peers;
    bool opened[kMaxWorld];
    bool connected = false;
    uint32_t epoch = 0, pool_waited = 0;
    // ranks on different GPUs: the pool goes to the peers on a second stream, beside the splice, which waits per parent
    bool split_ok = false;               // every peer is on another GPU (a spinning splice on a shared GPU would starve the peer)
    bool split_now = false;              // this generation's pool is being pushed that way
    bool push_outstanding = false;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_selected = nullptr, ev_pushed = nullptr;
    uint32_t* d_bounds = nullptr;        // [kMaxWorld + 1]
    uint32_t* d_push_counters = nullptr; // [kMaxWorld]
    uint32_t pci = 0;
    // device state owned by the shard
    uint32_t* d_counters = nullptr;      // [0] values push, [1] pool push
    int* d_err = nullptr;
    uint32_t* d_sel = nullptr;           // [max_pool] selected global indexes, ascending
    uint32_t* d_sel_ids = nullptr;       // [max_pool]
    Scratch s_local_rows;
    // pinned host staging, so that what the host mirror needs of a generation (local GEBVs, the picked pairs, the
    // selected parents' ids) comes down while the device carries on: [bvs: max_total f64 | picks: 2 max_total u32 | sel ids: max_pool u32]
    char* h_pin = nullptr;
    // ids on their way up: two pinned buffers in turn (a copy from pageable memory would make cudaMemcpyAsync wait for
    // everything queued before it, i.e. for the previous generation's splice and this generation's GEBVs)
    uint32_t* h_ids[2] = {nullptr, nullptr};
    cudaEvent_t ev_ids[2] = {nullptr, nullptr};
    bool ids_in_flight[2] = {false, false};
    // host copies (ids up; GEBVs and picks down) ride a stream of their own, forked from and joined to the context stream
    // by events: on the context stream each would sit between two kernels for the 10 - 30 us it takes
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_fork[3] = {nullptr, nullptr, nullptr};
    uint32_t crosses_since_select = 0;     // only the first cross of a selection may start its pair draws early (ev_fork[0] is its licence)
    bool picks_scratch_moved = false;
    uint64_t launches_at_select = 0;       // ... and only when nothing else was launched on this context in between (shared scratch)
    cudaEvent_t ev_bv = nullptr, ev_picks = nullptr, ev_flags = nullptr, ev_sel_ids = nullptr;
    int* h_flags = nullptr;              // pinned: [0] the shard's exchange flag, [1] the context's splice flag, as of the last poll
    bool poll_pending = false;
    bool bv_staged = false, picks_staged = false;
    uint32_t staged_local = 0, staged_offspring = 0;
    // optional phase timing: a ring of CUDA events, one set per generation (epoch)
    bool timing = false;
    cudaEvent_t ev[kTimingRing][kTimingEvents];
    bool ev_made = false;
    // last selection
    uint64_t n_total = 0;
    uint32_t top_n = 0;
    bool have_selection = false;
};

static void mark(gsb200_shard* s, uint32_t epoch, uint32_t which) {
    if (s->timing) cudaEventRecord(s->ev[epoch % kTimingRing][which], s->ctx->stream);
}

extern "C" {

void gsb200_shard_range(uint64_t n_total, uint32_t rank, uint32_t world, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = n_total / world, extra = n_total % world;
    const uint64_t a = rank * base + std::min<uint64_t>(rank, extra);
    if (lo) *lo = a;
    if (hi) *hi = a + base + (rank < extra ? 1 : 0);
}

int gsb200_shard_create(gsb200_ctx* ctx, uint32_t rank, uint32_t world, uint64_t max_total, uint32_t max_pool, gsb200_shard** out) {
    GSB_REQUIRE(ctx != nullptr && out != nullptr, GSB200_ERR_ARG, "gsb200_shard_create: null argument");
    *out = nullptr;
    GSB_REQUIRE(world >= 1 && world <= kMaxWorld && rank < world, GSB200_ERR_ARG, "gsb200_shard_create: rank %u of %u (at most %u ranks)", rank, world, kMaxWorld);
    GSB_REQUIRE(max_total >= 1 && max_total < 0xffffffffull && max_pool >= 1, GSB200_ERR_ARG, "gsb200_shard_create: bad sizes");
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    gsb200_shard* s = new gsb200_shard();
    s->ctx = ctx; s->rank = rank; s->world = world; s->max_total = max_total; s->max_pool = max_pool;
    s->row_words = ctx->row_words();
    memset(&s->peers, 0, sizeof s->peers);
    memset(s->opened, 0, sizeof s->opened);
    s->off_ids = align_up(kOffBv + max_total * sizeof(double), 256);
    s->off_pool = align_up(s->off_ids + max_total * sizeof(uint32_t), 256);
    s->slab_bytes = s->off_pool + (size_t) max_pool * s->row_words * sizeof(uint32_t);
    cudaError_t e = cudaMalloc((void**) &s->slab, s->slab_bytes);
    if (e == cudaSuccess) e = cudaMemset(s->slab, 0, kOffBv);
    if (e == cudaSuccess) e = cudaMalloc((void**) &s->d_counters, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->d_counters, 0, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**) &s->d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(s->d_err, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**) &s->d_sel, 2 * (size_t) max_pool * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void**) &s->h_pin, max_total * 16 + (size_t) max_pool * 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking);
    for (int k = 0; k < 3 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&s->ev_fork[k], cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaHostAlloc((void**) &s->h_ids[k], max_total * sizeof(uint32_t), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_ids[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_bv, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_picks, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_sel_ids, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_flags, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void**) &s->h_flags, 2 * sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void**) &s->d_bounds, (kMaxWorld + 1 + kMaxWorld) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->d_bounds, 0, (kMaxWorld + 1 + kMaxWorld) * sizeof(uint32_t));
    if (e == cudaSuccess) {
        int lo_pri = 0, hi_pri = 0;
        cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
        e = cudaStreamCreateWithPriority(&s->side, cudaStreamNonBlocking, hi_pri);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_selected, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_pushed, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        int dom = 0, bus = 0, dev = 0;
        cudaDeviceGetAttribute(&dom, cudaDevAttrPciDomainId, ctx->device);
        cudaDeviceGetAttribute(&bus, cudaDevAttrPciBusId, ctx->device);
        cudaDeviceGetAttribute(&dev, cudaDevAttrPciDeviceId, ctx->device);
        s->pci = ((uint32_t) dom << 16) | ((uint32_t) bus << 8) | (uint32_t) dev | 0x80000000u;
    }
    if (e != cudaSuccess) {
        set_error("gsb200_shard_create: allocating %zu bytes of exchange buffers failed: %s", s->slab_bytes, cudaGetErrorString(e));
        gsb200_shard_destroy(s);
        return e == cudaErrorMemoryAllocation ? GSB200_ERR_OOM : GSB200_ERR_CUDA;
    }
    s->d_sel_ids = s->d_sel + max_pool;
    s->d_push_counters = s->d_bounds + kMaxWorld + 1;
    s->peers.row_u4 = (uint32_t) (s->row_words / 4);
    // Every scratch buffer a generation touches is sized here, for the largest generation: nothing is
    // allocated (and, above all, nothing is cudaFree'd — which waits for the whole device, including a
    // peer shard of this process that is spinning on OUR next push) while generations are in flight.
    s->s_local_rows.stream = ctx->stream;
    int st = ctx->s_select.ensure(select_scratch_bytes((uint32_t) max_total));
    if (st == GSB200_OK) st = s->s_local_rows.ensure(max_total * sizeof(uint32_t));
    if (st == GSB200_OK) st = ctx->s_eval[2].ensure(max_total * 8 * sizeof(long long) + (size_t) (ctx->plane_words / 16 + 1) * sizeof(unsigned int));
    if (st == GSB200_OK) st = ctx->s_rows[1].ensure(2 * max_total * sizeof(uint32_t));
    if (st == GSB200_OK) st = ctx->s_rows[2].ensure(max_total * sizeof(uint32_t));
    if (st == GSB200_OK) st = ctx->s_rows[3].ensure(4 * max_total * sizeof(uint32_t));
    if (st != GSB200_OK) { gsb200_shard_destroy(s); return st; }
    s->peers.base[rank] = s->slab;
    s->connected = world == 1;
    *out = s;
    return GSB200_OK;
}

void gsb200_shard_destroy(gsb200_shard* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
    if (s->copy) { cudaStreamSynchronize(s->copy); cudaStreamDestroy(s->copy); }
    for (int k = 0; k < 3; ++k) if (s->ev_fork[k]) cudaEventDestroy(s->ev_fork[k]);
    if (s->ev_selected) cudaEventDestroy(s->ev_selected);
    if (s->ev_pushed) cudaEventDestroy(s->ev_pushed);
    if (s->d_bounds) cudaFree(s->d_bounds);
    for (uint32_t r = 0; r < kMaxWorld; ++r) if (s->opened[r]) cudaIpcCloseMemHandle(s->peers.base[r]);
    s->s_local_rows.release();
    if (s->ev_made) for (auto& set : s->ev) for (auto& e : set) cudaEventDestroy(e);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    for (int k = 0; k < 2; ++k) { if (s->h_ids[k]) cudaFreeHost(s->h_ids[k]); if (s->ev_ids[k]) cudaEventDestroy(s->ev_ids[k]); }
    if (s->ev_bv) cudaEventDestroy(s->ev_bv);
    if (s->ev_picks) cudaEventDestroy(s->ev_picks);
    if (s->ev_sel_ids) cudaEventDestroy(s->ev_sel_ids);
    if (s->ev_flags) cudaEventDestroy(s->ev_flags);
    if (s->h_flags) cudaFreeHost(s->h_flags);
    if (s->slab) cudaFree(s->slab);
    if (s->d_counters) cudaFree(s->d_counters);
    if (s->d_err) cudaFree(s->d_err);
    if (s->d_sel) cudaFree(s->d_sel);
    delete s;
}

int gsb200_shard_handle(gsb200_shard* s, void* handle) {
    GSB_REQUIRE(s != nullptr && handle != nullptr, GSB200_ERR_ARG, "gsb200_shard_handle: null argument");
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    ShardHandle h;
    memset(&h, 0, sizeof h);
    h.magic = kHandleMagic; h.rank = s->rank; h.world = s->world; h.device = (uint32_t) s->ctx->device;
    h.pid = (int32_t) getpid(); h.max_pool = s->max_pool;
    h.slab_bytes = s->slab_bytes; h.row_bytes = s->row_words * sizeof(uint32_t); h.max_total = s->max_total;
    h.raw = s->slab;
    h.pci = s->pci;
    GSB_CUDA_TRY(cudaIpcGetMemHandle(&h.ipc, s->slab));
    memset(handle, 0, GSB200_SHARD_HANDLE_BYTES);
    memcpy(handle, &h, sizeof h);
    return GSB200_OK;
}

int gsb200_shard_connect(gsb200_shard* s, const void* handles) {
    GSB_REQUIRE(s != nullptr && handles != nullptr, GSB200_ERR_ARG, "gsb200_shard_connect: null argument");
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    bool same_gpu = false;
    for (uint32_t r = 0; r < s->world; ++r) {
        ShardHandle h;
        memcpy(&h, (const char*) handles + (size_t) r * GSB200_SHARD_HANDLE_BYTES, sizeof h);
        GSB_REQUIRE(h.magic == kHandleMagic && h.rank == r && h.world == s->world, GSB200_ERR_ARG,
                    "gsb200_shard_connect: entry %u is not the handle of rank %u of %u", r, r, s->world);
        GSB_REQUIRE(h.slab_bytes == s->slab_bytes && h.row_bytes == s->row_words * sizeof(uint32_t) && h.max_total == s->max_total
                    && h.max_pool == s->max_pool, GSB200_ERR_ARG, "gsb200_shard_connect: rank %u was created with different sizes", r);
        if (r == s->rank) continue;
        if (h.pci == s->pci || h.pci == 0) same_gpu = true;
        if (s->opened[r]) { cudaIpcCloseMemHandle(s->peers.base[r]); s->opened[r] = false; }
        if (h.pid == (int32_t) getpid()) {
            // a shard of this very process: its slab is addressable as it is (peer access between two devices is switched on here)
            if ((int) h.device != s->ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess((int) h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("gsb200_shard_connect: no peer access from device %d to device %u: %s", s->ctx->device, h.device, cudaGetErrorString(e));
                    return GSB200_ERR_CUDA;
                }
                cudaGetLastError();
            }
            s->peers.base[r] = (char*) h.raw;
        } else {
            void* p = nullptr;
            GSB_CUDA_TRY(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            s->peers.base[r] = (char*) p;
            s->opened[r] = true;
        }
    }
    s->connected = true;
    s->split_ok = s->world > 1 && !same_gpu;
    return GSB200_OK;
}

static int shard_select_impl(gsb200_shard* s, uint32_t n_local, const uint32_t* d_local_rows, const uint32_t* local_ids,
                             uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best, const char* what) {
    gsb200_ctx* ctx = s->ctx;
    s->bv_staged = false;
    GSB_REQUIRE(s->connected, GSB200_ERR_ARG, "%s: the shard is not connected to its peers yet", what);
    GSB_REQUIRE(n_total >= 1 && n_total <= s->max_total, GSB200_ERR_ARG, "%s: generation of %llu exceeds the %llu the shard was created for",
                what, (unsigned long long) n_total, (unsigned long long) s->max_total);
    GSB_REQUIRE(top_n >= 1, GSB200_ERR_ARG, "%s: top_n must be positive", what);
    if (top_n > n_total) top_n = (uint32_t) n_total;            // "move them all" (core.c:9478-9484)
    GSB_REQUIRE(top_n <= s->max_pool, GSB200_ERR_ARG, "%s: top_n %u exceeds the pool of %u the shard was created for", what, top_n, s->max_pool);
    GSB_REQUIRE(ctx->row_words() == s->row_words, GSB200_ERR_ARG, "%s: the store changed its width since the shard was created", what);
    uint64_t lo, hi;
    gsb200_shard_range(n_total, s->rank, s->world, &lo, &hi);
    GSB_REQUIRE(n_local == hi - lo, GSB200_ERR_ARG, "%s: rank %u of %u holds individuals %llu..%llu of %llu, not %u", what, s->rank, s->world,
                (unsigned long long) lo, (unsigned long long) hi, (unsigned long long) n_total, n_local);
    GSB_REQUIRE(n_local == 0 || d_local_rows, GSB200_ERR_ARG, "%s: null row array", what);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    GSB_TRY(ctx->s_select.ensure(select_scratch_bytes((uint32_t) n_total)));
    const uint32_t prev = s->epoch;
    const uint32_t epoch = ++s->epoch;
    const volatile uint32_t* my_flags = reinterpret_cast<const volatile uint32_t*>(s->slab + kOffFlags);
    if (s->world > 1 && s->pool_waited < prev) {
        // nobody crossed from the previous selection here: make sure every peer is done reading the GEBV vector
        shard_wait_kernel<<<1, 32, 0, ctx->stream>>>(my_flags + kMaxWorld, s->world, prev, s->d_err);
        ++ctx->launches;
        s->pool_waited = prev;
    }
    if (s->push_outstanding) {             // the previous generation's remote push reads the selection state and the pool
        GSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_pushed, 0));
        s->push_outstanding = false;
    }
    double* all_bv = reinterpret_cast<double*>(s->slab + kOffBv);
    const int with_ids = local_ids != nullptr;
    // "everything queued so far is done": what the copy stream waits for before it touches anything the previous generation used
    GSB_CUDA_TRY(cudaEventRecord(s->ev_fork[0], ctx->stream));
    s->crosses_since_select = 0;
    if (with_ids && n_local) {
        // behind everything queued so far (the previous generation's readers of the id column), beside this generation's GEBVs
        const int b = (int) (epoch & 1u);
        if (s->ids_in_flight[b]) GSB_CUDA_TRY(cudaEventSynchronize(s->ev_ids[b]));      // two generations ago: long done
        memcpy(s->h_ids[b], local_ids, (size_t) n_local * sizeof(uint32_t));
        GSB_CUDA_TRY(cudaStreamWaitEvent(s->copy, s->ev_fork[0], 0));
        GSB_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(s->slab + s->off_ids) + lo, s->h_ids[b], (size_t) n_local * sizeof(uint32_t),
                                     cudaMemcpyHostToDevice, s->copy));
        GSB_CUDA_TRY(cudaEventRecord(s->ev_ids[b], s->copy));
        s->ids_in_flight[b] = true;
    }
    mark(s, epoch, 0);
    if (n_local) GSB_TRY(gsb200_gebv_dev(ctx, n_local, d_local_rows, eff_index, all_bv + lo));
    if (d_local_rows == (const uint32_t*) s->s_local_rows.ptr) {
        // the host-row entry point: this rank's GEBVs start their way down now, behind nothing but the GEBV kernels
        GSB_CUDA_TRY(cudaEventRecord(s->ev_fork[1], ctx->stream));
        GSB_CUDA_TRY(cudaStreamWaitEvent(s->copy, s->ev_fork[1], 0));
        if (n_local) GSB_CUDA_TRY(cudaMemcpyAsync(s->h_pin, all_bv + lo, (size_t) n_local * sizeof(double), cudaMemcpyDeviceToHost, s->copy));
        GSB_CUDA_TRY(cudaEventRecord(s->ev_bv, s->copy));
        s->bv_staged = true;
        s->staged_local = n_local;
    }
    if (with_ids && n_local) GSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_ids[epoch & 1u], 0));      // the ids are up (sent while the GEBVs were computed)
    mark(s, epoch, 1);
    if (s->world > 1) {
        const unsigned grid = (unsigned) std::max<uint32_t>(1, std::min<uint32_t>((n_local + 255) / 256, (uint32_t) ctx->sm_count * 2));
        shard_push_values_kernel<<<grid, 256, 0, ctx->stream>>>(s->peers, s->world, s->rank, s->off_ids, (uint32_t) lo, n_local, with_ids,
                                                                s->d_counters, epoch);
        shard_wait_kernel<<<1, 32, 0, ctx->stream>>>(my_flags, s->world, epoch, s->d_err);
        ctx->launches += 2;
    }
    mark(s, epoch, 2);
    GSB_TRY(select_top_n(ctx, all_bv, (uint32_t) n_total, top_n, low_is_best, ctx->s_select.ptr, s->d_sel, (uint32_t) lo, (uint32_t) hi));
    mark(s, epoch, 3);
    // Ranks on different GPUs and enough to send for it to matter (>= 24 MB: the splice then sorts its gametes, own parents first): my rows go into my own pool buffer here, and to
    // the peers on the second stream while the splice is already at work on the parents it has (shard_cross_impl).
    const size_t remote_bytes = (size_t) top_n * s->row_words * sizeof(uint32_t) / s->world * (s->world - 1);
    // MEASURED AND LEFT OFF (GSB200_SHARD_SPLIT_PUSH=1: from 64 MB to send on, =2: always — the two-GPU test forces it, =3: the
    // same kernels on one stream): the push phase shrinks as intended (N = 8: 0.145 -> 0.026 ms) but one peer at a time the push
    // runs at 0.5 TB/s where the all-peers-at-once kernel reaches 0.6 - 0.8, the splice ends up waiting for the last peers'
    // rows (0.450 -> 0.603 ms) and the step is 0.03 - 0.05 ms SLOWER on 2, 4 and 8 GPUs (profiles/r2_split_push.txt).
    static const int split_mode = [] { const char* v = getenv("GSB200_SHARD_SPLIT_PUSH"); return v ? atoi(v) : 0; }();
    s->split_now = s->split_ok && (((split_mode == 1 || split_mode == 3) && remote_bytes >= ((size_t) 64 << 20)) || split_mode == 2);
    if (s->split_now) {
        shard_owner_bounds_kernel<<<1, 32, 0, ctx->stream>>>(s->d_sel, top_n, n_total, s->world, s->d_bounds);
        shard_push_pool_local_kernel<<<(unsigned) ctx->sm_count * 2, 256, 0, ctx->stream>>>(
            s->peers, s->rank, s->off_ids, s->off_pool, ctx->d_store, s->row_words, d_local_rows, (uint32_t) lo,
            (const SelectState*) ctx->s_select.ptr, s->d_sel, top_n, s->d_sel_ids, with_ids, s->d_counters + 1, epoch);
        if (split_mode == 3) {          // measurement: the same kernels, one after the other on one stream
            shard_push_pool_remote_kernel<<<(unsigned) ctx->sm_count, 32, 0, ctx->stream>>>(
                s->peers, s->world, s->rank, s->off_pool, s->d_bounds, s->d_push_counters, epoch);
        } else {
            GSB_CUDA_TRY(cudaEventRecord(s->ev_selected, ctx->stream));
            GSB_CUDA_TRY(cudaStreamWaitEvent(s->side, s->ev_selected, 0));
            shard_push_pool_remote_kernel<<<(unsigned) ctx->sm_count, 32, 0, s->side>>>(
                s->peers, s->world, s->rank, s->off_pool, s->d_bounds, s->d_push_counters, epoch);
            GSB_CUDA_TRY(cudaEventRecord(s->ev_pushed, s->side));
            s->push_outstanding = true;
        }
        ctx->launches += 3;
    } else {
        shard_push_pool_kernel<<<(unsigned) ctx->sm_count * 4, 256, 0, ctx->stream>>>(
            s->peers, s->world, s->rank, s->off_ids, s->off_pool, ctx->d_store, s->row_words, d_local_rows, (uint32_t) lo,
            (const SelectState*) ctx->s_select.ptr, s->d_sel, top_n, s->d_sel_ids, with_ids, s->d_counters + 1, epoch);
        ++ctx->launches;
    }
    GSB_CUDA_TRY(cudaGetLastError());
    if (s->bv_staged) GSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_bv, 0));      // joined: nothing later overwrites GEBVs still on their way down
    mark(s, epoch, 4);
    s->launches_at_select = ctx->launches;
    s->n_total = n_total;
    s->top_n = top_n;
    s->have_selection = true;
    return GSB200_OK;
}

int gsb200_shard_select_dev(gsb200_shard* s, uint32_t n_local, const uint32_t* d_local_rows, const uint32_t* local_ids,
                            uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    return shard_select_impl(s, n_local, d_local_rows, local_ids, n_total, eff_index, top_n, low_is_best, "gsb200_shard_select_dev");
}

int gsb200_shard_select_run(gsb200_shard* s, uint32_t n_local, const uint32_t* local_rows, uint32_t first_local_row, const uint32_t* local_ids,
                            uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    const char* what = "gsb200_shard_select";
    gsb200_ctx* ctx = s->ctx;
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    // the rows are read again when the pool is pushed: they get a buffer of their own
    uint32_t* d_rows;
    if (local_rows) {
        GSB_TRY(check_rows(ctx, n_local, local_rows, what));
        GSB_TRY(upload_u32(ctx, s->s_local_rows, local_rows, n_local, &d_rows));
    } else {
        GSB_REQUIRE((uint64_t) first_local_row + n_local <= ctx->capacity, GSB200_ERR_ARG, "%s: rows %u.. beyond the store capacity", what, first_local_row);
        GSB_TRY(s->s_local_rows.ensure(std::max<size_t>(n_local, 1) * sizeof(uint32_t)));
        d_rows = (uint32_t*) s->s_local_rows.ptr;
        if (n_local) shard_iota_kernel<<<(n_local + 255) / 256, 256, 0, ctx->stream>>>(d_rows, first_local_row, n_local);
    }
    return shard_select_impl(s, n_local, d_rows, local_ids, n_total, eff_index, top_n, low_is_best, what);
}

int gsb200_shard_select(gsb200_shard* s, uint32_t n_local, const uint32_t* local_rows, const uint32_t* local_ids,
                        uint64_t n_total, uint32_t eff_index, uint32_t top_n, int low_is_best) {
    GSB_REQUIRE(n_local == 0 || local_rows != nullptr, GSB200_ERR_ARG, "gsb200_shard_select: null row array");
    return gsb200_shard_select_run(s, n_local, local_rows, 0, local_ids, n_total, eff_index, top_n, low_is_best);
}

static int shard_cross_impl(gsb200_shard* s, uint32_t n, uint32_t map, const uint32_t* d_out_rows, uint32_t first_out_row,
                            const gsb200_draws* draws, uint32_t* d_picks, bool stage_picks, const char* what) {
    GSB_REQUIRE(s->have_selection, GSB200_ERR_ARG, "%s: nothing has been selected yet", what);
    gsb200_ctx* ctx = s->ctx;
    const uint64_t launches_before = ctx->launches;
    GSB_REQUIRE(ctx->row_words() == s->row_words, GSB200_ERR_ARG, "%s: the store changed its width since the shard was created", what);
    GSB_REQUIRE(!stage_picks || n <= s->max_total, GSB200_ERR_ARG, "%s: %u offspring exceed the %llu the shard was created for", what, n,
                (unsigned long long) s->max_total);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const volatile uint32_t* pool_flags = reinterpret_cast<const volatile uint32_t*>(s->slab + kOffFlags) + kMaxWorld;
    // a pool pushed in two parts is waited for parent by parent inside the splice (and by everybody once more behind it);
    // otherwise here, as a whole
    const bool lazy = s->split_now && s->pool_waited < s->epoch;
    if (s->world > 1 && s->pool_waited < s->epoch && !lazy) {
        shard_wait_kernel<<<1, 32, 0, ctx->stream>>>(pool_flags, s->world, s->epoch, s->d_err);
        ++ctx->launches;
        s->pool_waited = s->epoch;
    }
    PoolWait pw;
    pw.flags = pool_flags; pw.bounds = s->d_bounds; pw.epoch = s->epoch; pw.world = s->world; pw.rank = s->rank; pw.err = s->d_err;
    mark(s, s->epoch, 5);
    uint32_t* h_picks = nullptr;
    static const bool early_on = [] { const char* v = getenv("GSB200_SHARD_EARLY_PAIRS"); return !v || atoi(v) != 0; }();      // =0: pairs in line (A/B)
    const bool early_ok = early_on && s->crosses_since_select++ == 0 && !s->picks_scratch_moved && launches_before == s->launches_at_select;
    s->picks_scratch_moved = false;
    if (stage_picks) {
        h_picks = reinterpret_cast<uint32_t*>(s->h_pin + s->max_total * 8);
        GSB_CUDA_TRY(cudaMemcpyAsync(s->h_pin + s->max_total * 16, s->d_sel_ids, (size_t) s->top_n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaEventRecord(s->ev_sel_ids, ctx->stream));      // (the picks may be down long before: they do not wait for the selection)
    }
    GSB_TRY(pool_random_cross(ctx, reinterpret_cast<const uint32_t*>(s->slab + s->off_pool), s->top_n, n, map, d_out_rows, first_out_row,
                              draws, d_picks, what, h_picks, s->ev_picks, lazy ? &pw : nullptr,
                              early_ok ? s->copy : nullptr, s->ev_fork[0], s->ev_fork[2]));
    if (lazy) {
        // every peer's rows have been seen before anything that follows on this stream: the GEBVs of the next generation
        // are pushed only after this, which is what lets the peers reuse their buffers without a barrier
        shard_wait_kernel<<<1, 32, 0, ctx->stream>>>(pool_flags, s->world, s->epoch, s->d_err);
        ++ctx->launches;
        s->pool_waited = s->epoch;
    }
    s->picks_staged = stage_picks;
    s->staged_offspring = n;
    mark(s, s->epoch, 6);
    return GSB200_OK;
}

int gsb200_shard_cross_dev(gsb200_shard* s, uint32_t n, uint32_t map, const uint32_t* d_out_rows, uint32_t first_out_row,
                           const gsb200_draws* draws, uint32_t* d_picks) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    return shard_cross_impl(s, n, map, d_out_rows, first_out_row, draws, d_picks, false, "gsb200_shard_cross_dev");
}

int gsb200_shard_cross_begin(gsb200_shard* s, uint32_t n, uint32_t map, const uint32_t* out_rows, const gsb200_draws* draws, int want_picks) {
    GSB_REQUIRE(n == 0 || out_rows != nullptr, GSB200_ERR_ARG, "gsb200_shard_cross: null row array");
    return gsb200_shard_cross_begin_run(s, n, map, out_rows, 0, draws, want_picks);
}

int gsb200_shard_cross_begin_run(gsb200_shard* s, uint32_t n, uint32_t map, const uint32_t* out_rows, uint32_t first_out_row,
                                 const gsb200_draws* draws, int want_picks) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    if (n == 0) return GSB200_OK;
    gsb200_ctx* ctx = s->ctx;
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t* d_out = nullptr;
    if (out_rows) {
        GSB_TRY(check_rows(ctx, n, out_rows, "gsb200_shard_cross"));
        GSB_TRY(upload_u32(ctx, ctx->s_rows[2], out_rows, n, &d_out));
    }
    uint32_t* d_picks = nullptr;
    if (want_picks) {
        const void* before = ctx->s_rows[1].ptr;
        GSB_TRY(ctx->s_rows[1].ensure((size_t) 2 * n * sizeof(uint32_t)));
        d_picks = (uint32_t*) ctx->s_rows[1].ptr;
        s->picks_scratch_moved = before != (const void*) d_picks;
    }
    return shard_cross_impl(s, n, map, d_out, first_out_row, draws, d_picks, want_picks != 0, "gsb200_shard_cross");
}

int gsb200_shard_cross_picks(gsb200_shard* s, const uint32_t** picked1, const uint32_t** picked2, const uint32_t** selected_ids) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    GSB_REQUIRE(s->picks_staged, GSB200_ERR_ARG, "gsb200_shard_cross_picks: the last cross was not asked for its picks");
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    GSB_CUDA_TRY(cudaEventSynchronize(s->ev_picks));
    GSB_CUDA_TRY(cudaEventSynchronize(s->ev_sel_ids));
    const uint32_t* p = reinterpret_cast<const uint32_t*>(s->h_pin + s->max_total * 8);
    if (picked1) *picked1 = p;
    if (picked2) *picked2 = p + s->staged_offspring;
    if (selected_ids) *selected_ids = reinterpret_cast<const uint32_t*>(s->h_pin + s->max_total * 16);
    return GSB200_OK;
}

int gsb200_shard_poll(gsb200_shard* s) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    gsb200_ctx* ctx = s->ctx;
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (s->poll_pending && cudaEventQuery(s->ev_flags) == cudaSuccess) {
        s->poll_pending = false;
        if (s->h_flags[0] || s->h_flags[1]) return gsb200_shard_finish(s);      // something went wrong: the blocking path words it (and clears the flags)
    }
    cudaGetLastError();
    if (!s->poll_pending) {
        GSB_CUDA_TRY(cudaMemcpyAsync(s->h_flags, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaMemcpyAsync(s->h_flags + 1, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaEventRecord(s->ev_flags, ctx->stream));
        s->poll_pending = true;
    }
    return GSB200_OK;
}

int gsb200_shard_finish(gsb200_shard* s) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    GSB_TRY(gsb200_cross_dev_status(s->ctx));           // synchronises; a plan overflow is reported, not silently kept
    return gsb200_shard_status(s);
}

int gsb200_shard_cross(gsb200_shard* s, uint32_t n, uint32_t map, const uint32_t* out_rows, const gsb200_draws* draws,
                       uint32_t* picked1, uint32_t* picked2) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    if (n == 0) return GSB200_OK;
    const int want = picked1 && picked2;
    GSB_TRY(gsb200_shard_cross_begin(s, n, map, out_rows, draws, want));
    if (want) {
        const uint32_t *p1, *p2;
        GSB_TRY(gsb200_shard_cross_picks(s, &p1, &p2, nullptr));
        memcpy(picked1, p1, (size_t) n * sizeof(uint32_t));
        memcpy(picked2, p2, (size_t) n * sizeof(uint32_t));
    }
    return gsb200_shard_finish(s);
}

int gsb200_shard_selected(gsb200_shard* s, uint32_t* global_indexes, uint32_t* ids, double* all_bvs) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    GSB_REQUIRE(s->have_selection, GSB200_ERR_ARG, "gsb200_shard_selected: nothing has been selected yet");
    gsb200_ctx* ctx = s->ctx;
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (global_indexes) GSB_CUDA_TRY(cudaMemcpyAsync(global_indexes, s->d_sel, (size_t) s->top_n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (ids) GSB_CUDA_TRY(cudaMemcpyAsync(ids, s->d_sel_ids, (size_t) s->top_n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (all_bvs) GSB_CUDA_TRY(cudaMemcpyAsync(all_bvs, s->slab + kOffBv, (size_t) s->n_total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return gsb200_shard_status(s);
}

int gsb200_shard_set_timing(gsb200_shard* s, int enabled) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (enabled && !s->ev_made) {
        for (auto& set : s->ev) for (auto& e : set) GSB_CUDA_TRY(cudaEventCreate(&e));
        s->ev_made = true;
    }
    s->timing = enabled != 0;
    return GSB200_OK;
}

uint32_t gsb200_shard_epoch(const gsb200_shard* s) { return s ? s->epoch : 0; }

int gsb200_shard_phase_ms(gsb200_shard* s, uint32_t epoch, float* ms) {
    GSB_REQUIRE(s != nullptr && ms != nullptr, GSB200_ERR_ARG, "gsb200_shard_phase_ms: null argument");
    GSB_REQUIRE(s->ev_made && epoch >= 1 && epoch <= s->epoch && s->epoch - epoch < kTimingRing, GSB200_ERR_ARG,
                "gsb200_shard_phase_ms: generation %u was not timed or is no longer kept", epoch);
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    GSB_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    cudaEvent_t* e = s->ev[epoch % kTimingRing];
    for (uint32_t k = 0; k + 1 < kTimingEvents; ++k) GSB_CUDA_TRY(cudaEventElapsedTime(&ms[k], e[k], e[k + 1]));
    return GSB200_OK;
}

int gsb200_shard_local_bvs(gsb200_shard* s, uint32_t n_local, double* out) {
    GSB_REQUIRE(s != nullptr && out != nullptr, GSB200_ERR_ARG, "gsb200_shard_local_bvs: null argument");
    GSB_REQUIRE(s->have_selection, GSB200_ERR_ARG, "gsb200_shard_local_bvs: nothing has been evaluated yet");
    uint64_t lo, hi;
    gsb200_shard_range(s->n_total, s->rank, s->world, &lo, &hi);
    GSB_REQUIRE(n_local == hi - lo, GSB200_ERR_ARG, "gsb200_shard_local_bvs: this rank evaluated %llu individuals, not %u",
                (unsigned long long) (hi - lo), n_local);
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (s->bv_staged && s->staged_local == n_local) {           // already on their way since the selection was enqueued
        GSB_CUDA_TRY(cudaEventSynchronize(s->ev_bv));
        memcpy(out, s->h_pin, (size_t) n_local * sizeof(double));
        return GSB200_OK;
    }
    if (n_local) GSB_CUDA_TRY(cudaMemcpyAsync(out, reinterpret_cast<const double*>(s->slab + kOffBv) + lo, (size_t) n_local * sizeof(double),
                                              cudaMemcpyDeviceToHost, s->ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    return GSB200_OK;
}

uint64_t gsb200_shard_generation_size(const gsb200_shard* s) { return s && s->have_selection ? s->n_total : 0; }
uint32_t gsb200_shard_pool_size(const gsb200_shard* s) { return s && s->have_selection ? s->top_n : 0; }
void* gsb200_shard_pool_device_ptr(gsb200_shard* s) { return s ? (void*) (s->slab + s->off_pool) : nullptr; }

int gsb200_shard_status(gsb200_shard* s) {
    GSB_REQUIRE(s != nullptr, GSB200_ERR_ARG, "null shard");
    GSB_CUDA_TRY(cudaSetDevice(s->ctx->device));
    int err = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    GSB_REQUIRE(err == 0, GSB200_ERR_CUDA, "shard %u of %u: a peer did not deliver its part of the exchange in time", s->rank, s->world);
    return GSB200_OK;
}

}  // extern "C"
