// local_tc5.cu — local (marker-block) GEBVs for MANY effect sets on the 5th-generation tensor cores:
// the dense-contraction path of BASELINE.json configs[4] (gsc_calculate_utility_local_bvs,
// core.c:9979-10090, computed for up to 10 effect sets in one pass over the genotypes).
//
// The contraction: (strands x markers) bits  x  (markers x 8 * sets) fixed-point digits, int32
// accumulation per marker block, recombined in fp64 per (strand, block, set).  With N = 8 * sets columns
// one A operand serves every effect set.  What the instructions cost decided the structure
// (profiles/r2_umma_i8_probe2.txt, r2_commit_probe.txt, r2_tmem_bw_probe.txt): a tcgen05.mma.kind::i8 of
// M = 128, K = 32 takes 47 cycles up to N = 96 (then 128 N / 256) — ten sets cost what one does — but a
// tcgen05.commit stalls the issuing thread ~300-400 cycles, so everything that is recycled behind the MMAs
// is recycled in UNITS of up to 20 steps with ONE commit per unit.
//
// Work item = (tile of 64 individuals = 128 strand rows = the 128 TMEM lanes) x (range of whole blocks).
// A CTA walks its items; inside an item everything streams:
//   * A-loader warp: cp.async, 8 lanes per strand row and 128-byte line (32 marker words), into a ring
//     of 3 lines in shared memory — the rows of a tile are 2 x plane_words apart, so a lane per row (the
//     first version) made every warp request touch 32 different lines;
//   * a unit = the steps of one 16-word half line (cut at 20 steps); 4 expander warps: thread t reads ITS
//     row's 64 bytes of the half line, turns each word into 8 operand registers `w & (0x01010101 << s)`
//     (byte b = bit 8b + s at weight 2^s; the weight is divided out of the digits on the host) and writes them
//     to its TMEM lane with tcgen05.st.32x32b into one of two operand stages — one K = 32 step per word, so a
//     step is 32 CONTIGUOUS markers and a block boundary costs at most one extra step (the word that
//     straddles it, B digits zeroed outside the block);
//   * B-loader thread: a unit's step records and digits in one cp.async.bulk into a ring of stages; B is
//     re-read from L2 for every tile (2 560 bytes per step at 10 sets), which keeps every block's sum in
//     ONE accumulator — the first version kept a marker segment's B resident instead and paid for it with
//     fp64 atomics on every block that a segment boundary cut;
//   * issuer thread: per step one tcgen05.mma.kind::i8, M = 128, N = 8 * sets (padded to 16), K = 32, A from
//     TMEM, B through a shared-memory descriptor; three accumulators in rotation, one per block in flight;
//     a commit when a block ends (its sums are ready) and one per unit (operand and B stage are free);
//   * 4 epilogue warps: tcgen05.ld of a finished block's 8 * sets digit sums per strand, exact integer
//     recombination, scale + base in fp64, then four blocks at a time are transposed through shared memory
//     so that a warp store writes 32 contiguous bytes per strand row instead of 8; blocks without markers
//     are written as 0 here (no memset of the outputs).
//
// Every mbarrier wait is bounded (tc5_common.cuh): a protocol bug raises the context's flag instead of
// hanging the GPU.
#include "gsb200_internal.cuh"
#include "tc5_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace gsb;
using namespace gsb::tc5;

namespace gsb {

namespace {

constexpr uint32_t kEpiWarps = 8;                // two per TMEM lane quarter: even / odd effect sets
constexpr uint32_t kWarpIssuer = 4 + kEpiWarps, kWarpALoader = kWarpIssuer + 2, kWarpBLoader = kWarpIssuer + 3;      // two issuer warps, alternating units
constexpr uint32_t kThreads = (kWarpBLoader + 1) * 32;           // 4 expander + 8 epilogue warps, 2 issuers, A loader, B loader
constexpr uint32_t kALines = 2;                  // ring of 128-byte lines (32 marker words x 128 strand rows) in shared memory
constexpr uint32_t kARowBytes = 144;             // 128 + 16: the expanders' 16-byte reads of 8 different rows hit different banks
constexpr uint32_t kALineBytes = 128 * kARowBytes;
constexpr uint32_t kUnitWords = 16, kUnitSteps = 17;      // 17: three B stages of ten effect sets fit beside the rest
constexpr uint32_t kStageCols = 8 * kUnitWords;  // an operand stage in TMEM: 16 words x 8 columns
constexpr uint32_t kACol0 = 256;                 // two operand stages: columns 256..511
constexpr uint32_t kSlots = 3, kSlotCols = 80;   // three accumulators: columns 0..239
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kUnitHead = 256;              // in front of a unit's digits: n_segments, its segment records (zero-padded)
constexpr uint32_t kStageSlack = 256;             // an odd number of sets: the mma's 16-column granule reads 256 bytes past the last step
constexpr uint32_t kMaxBStages = 4;
constexpr uint32_t kOutRow = 5;                  // staging: 4 blocks per strand row + 1 double of padding
constexpr uint32_t kSmemBudget = 227 * 1024 - 2048;
// segment record (consecutive words of one block within a unit): first word within the half line (4 bits) | steps << 4 |
// 512 the block starts here | 1024 the block ends here
constexpr uint32_t kSegFirst = 512u, kSegLast = 1024u;
static_assert(kUnitSteps <= 20, "the MMA chain of the issuer is written out for up to 20 steps");
static_assert(kLocalTc5MaxSets * 8 <= kSlotCols, "an accumulator holds 8 columns per effect set");

struct Tc5Range {          // the whole blocks [blk_lo, blk_hi) = runs [run_lo, run_hi) of non-empty blocks
    uint32_t unit_lo, n_units;
    uint32_t run_lo, run_hi;
    uint32_t line_lo, n_lines;  // 128-byte lines of the strand rows its steps read
    uint32_t blk_lo, blk_hi;
};
struct Tc5Unit {
    uint32_t b_off16, b_bytes;  // the unit in the B stream (offset in 16-byte units): head + n_steps x 256 x sets
    uint32_t line_adv_half;     // lines to advance before the unit << 1 | which half of the line
    uint32_t n_ends;            // blocks that end inside the unit
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    return leader != 0;
}
// the kind::i8 MMA of tc5_common.cuh without the optional lane mask and with the B descriptor in halves (only the low
// one, which holds the start address, changes from step to step)
__device__ __forceinline__ void umma_i8_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 d;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], d, %4, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "r"(tmem_a), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const uint32_t* gmem_src) {
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" :: "r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}

// four marker words of this thread's strand row -> 4 K-steps of operand bytes in its TMEM lane
__device__ __forceinline__ void store_words(uint32_t taddr, const uint4 w) {
    uint32_t r[32];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const uint32_t m = 0x01010101u << s;
        r[s] = w.x & m; r[8 + s] = w.y & m; r[16 + s] = w.z & m; r[24 + s] = w.w & m;
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

struct LocalTc5Params {
    const uint4* b[kLocalTc5MaxSets];      // per set: [n_steps][16] uint4 = 256 bytes per step, UMMA K-major layout
    const double* base[kLocalTc5MaxSets];  // [n_blocks]
    double* out[kLocalTc5MaxSets];         // [(2n) x n_blocks]
    double scale[kLocalTc5MaxSets];
};

// The B stream: per unit a 256-byte head (n_steps, the step records) then [step][set][256 bytes of digits].
// One thread block per unit.
__global__ void local_tc5_pack_kernel(uint4* __restrict__ bstream, const Tc5Unit* __restrict__ units, uint32_t n_sets,
                                      const uint32_t* __restrict__ unit_step /* [n_units + 1] first step of every unit */,
                                      const uint32_t* __restrict__ heads /* [64 n_units] */, const LocalTc5Params P) {
    const Tc5Unit u = units[blockIdx.x];
    uint4* dst = bstream + u.b_off16;
    const uint32_t s0 = unit_step[blockIdx.x], total = u.b_bytes / 16;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        uint4 v;
        if (i < kUnitHead / 16) {
            const uint32_t* hp = heads + blockIdx.x * (kUnitHead / 4) + i * 4;
            v = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        } else {
            const uint32_t r0 = i - kUnitHead / 16, pos = r0 / (n_sets * 16), r = r0 % (n_sets * 16), q = r >> 4, k = r & 15;
            v = __ldg(P.b[q] + (size_t) (s0 + pos) * 16 + k);
        }
        dst[i] = v;
    }
}

__global__ void __launch_bounds__(kThreads, 1) local_tc5_kernel(
        const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
        const uint32_t* __restrict__ rows, uint32_t n, uint32_t n_blocks,
        const uint8_t* __restrict__ bstream, const Tc5Range* __restrict__ ranges, uint32_t n_ranges, const Tc5Unit* __restrict__ units,
        const uint32_t* __restrict__ runs /* block of every run */, uint32_t n_items, uint32_t n_sets, uint32_t n_bstages,
        const LocalTc5Params P, int* __restrict__ flag,
        uint32_t dbg /* GSB200_LOCAL_DEBUG bits, measurements only: 1 no operand stores, 2 no mma, 4 no epilogue work, 8 no row loads, 16 no B digits (heads only) */,
        long long* __restrict__ prof /* GSB200_LOCAL_PROF: CTA 0's cycles per role spent in each kind of wait, or null */) {
    extern __shared__ __align__(128) uint8_t s_dyn[];               // A lines | output staging | B stages
    __shared__ __align__(8) uint64_t s_afull[kALines], s_afree[kALines], s_tfull[2], s_issued[2], s_udone[kMaxBStages];
    __shared__ __align__(8) uint64_t s_bfull[kMaxBStages], s_dready[kSlots], s_dfree[kSlots];
    __shared__ const uint32_t* s_rowptr[128];
    __shared__ uint32_t s_tmem;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t i = 0; i < kALines; ++i) { mbar_init(&s_afull[i], 32); mbar_init(&s_afree[i], 4); }
        for (uint32_t i = 0; i < 2; ++i) { mbar_init(&s_tfull[i], 4); mbar_init(&s_issued[i], 1); }
        for (uint32_t i = 0; i < kMaxBStages; ++i) { mbar_init(&s_bfull[i], 1); mbar_init(&s_udone[i], 1); }
        for (uint32_t i = 0; i < kSlots; ++i) { mbar_init(&s_dready[i], 1); mbar_init(&s_dfree[i], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpIssuer) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t bar_afull = smem_u32(s_afull), bar_afree = smem_u32(s_afree), bar_tfull = smem_u32(s_tfull), bar_udone = smem_u32(s_udone), bar_issued = smem_u32(s_issued);
    const uint32_t bar_bfull = smem_u32(s_bfull), bar_dready = smem_u32(s_dready), bar_dfree = smem_u32(s_dfree);
    const uint32_t step_bytes = 256 * n_sets, stage_bytes = kUnitHead + kUnitSteps * step_bytes + kStageSlack;
    const uint32_t a_smem = smem_u32(s_dyn);
    const uint32_t sets_warp = (n_sets + 1) / 2;                    // effect sets per epilogue warp
    const uint32_t out_bytes_warp = sets_warp * 32 * kOutRow * 8;
    double* const out_stage0 = reinterpret_cast<double*>(s_dyn + kALines * kALineBytes);
    const uint32_t b_smem = a_smem + kALines * kALineBytes + kEpiWarps * out_bytes_warp;
    // M = 128 wants N % 16 == 0: an odd number of sets contracts 8 more columns against the bytes that follow
    // (the next step's first set, or whatever the stage held before) into accumulator columns nobody reads
    const uint32_t n_cols = (8 * n_sets + 15) & ~15u;
    const uint32_t idesc = idesc_i8(n_cols);
    // the unit barriers form a ring as long as the B ring (>= 2): unit U is committed to barrier U % n_bstages, and
    // both who wait for it (the expanders for unit U + 2's operand stage, the B loader for unit U + n_bstages' stage)
    // do so before the issuer can get a full ring further
    const uint32_t n_ring = n_bstages;
    bool ok = true;
    // prof[role * 4 + k]: k = 0 the role's whole loop, 1..3 its waits (see local_tc5_launch)
    const bool profiling = prof != nullptr && blockIdx.x == 0 && lane == 0;
    long long p_wait[3] = { 0, 0, 0 };
    const long long p_t0 = clock64();
    // whole-warp wait: ONE lane polls the barrier, the others park at the warp barrier.  With every lane of nine
    // waiting warps polling, the shared-memory pipe was so busy with phase checks that the issuer's record
    // loads took ~200 cycles each (profiles/r2_local_tc5.txt).
    auto wait_on = [&](uint32_t bar, uint32_t parity, int kind) {
        if (lane == 0) {
            if (profiling) { const long long t = clock64(); mbar_wait(bar, parity, ok); p_wait[kind] += clock64() - t; }
            else mbar_wait(bar, parity, ok);
        }
        __syncwarp();
    };
    auto prof_out = [&](uint32_t role) {
        if (profiling) { prof[role * 4] = clock64() - p_t0; prof[role * 4 + 1] = p_wait[0]; prof[role * 4 + 2] = p_wait[1]; prof[role * 4 + 3] = p_wait[2]; }
    };

    auto range_of = [&](uint32_t item) {
        const uint4* p = reinterpret_cast<const uint4*>(ranges + item % n_ranges);
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        Tc5Range r;
        r.unit_lo = a.x; r.n_units = a.y; r.run_lo = a.z; r.run_hi = a.w; r.line_lo = b.x; r.n_lines = b.y; r.blk_lo = b.z; r.blk_hi = b.w;
        return r;
    };

    if (warp < 4) {
        // ---------------- expanders: thread tid = strand row tid of the tile = TMEM lane tid
        uint32_t q_line = 0, q_unit = 0, ring_i = 0, ring_par = 0;       // ring_i = q_unit % n_ring, ring_par = (q_unit / n_ring) & 1
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Tc5Range rg = range_of(item);
            uint32_t lines_taken = 0, as = 0;
            auto next_line = [&]() {
                if (lines_taken && lane == 0) mbar_arrive(bar_afree + 8 * as);       // every lane's words of the old line were consumed by its stores
                as = q_line % kALines;
                wait_on(bar_afull + 8 * as, (q_line / kALines) & 1, 0);
                ++q_line; ++lines_taken;
            };
            uint32_t desc_next = __ldg(&units[rg.unit_lo].line_adv_half);
#pragma unroll 1
            for (uint32_t u = 0; u < rg.n_units; ++u, ++q_unit) {
                const uint32_t desc = desc_next;
                if (u + 1 < rg.n_units) desc_next = __ldg(&units[rg.unit_lo + u + 1].line_adv_half);      // an L2 round trip: asked for a unit ahead
#pragma unroll 1
                for (uint32_t k = desc >> 1; k; --k) next_line();
                const uint32_t src = a_smem + as * kALineBytes + tid * kARowBytes + (desc & 1u) * 64;
                uint4 w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = lds128(src + 16 * k);
                if (q_unit >= 2) {
                    // the operand stage was read by unit q_unit - 2: its barrier is two behind in the ring
                    const uint32_t back = ring_i >= 2 ? ring_i - 2 : ring_i + n_ring - 2;
                    const uint32_t par = ring_i >= 2 ? ring_par : ring_par ^ 1;
                    wait_on(bar_udone + 8 * back, par, 1);
                    tc_fence_after();
                }
                const uint32_t taddr = tmem + ((warp * 32) << 16) + kACol0 + kStageCols * (q_unit & 1);
                if (!(dbg & 1u)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) store_words(taddr + 32 * k, w[k]);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tfull + 8 * (q_unit & 1));
                if (++ring_i == n_ring) { ring_i = 0; ring_par ^= 1; }
            }
#pragma unroll 1
            while (lines_taken < rg.n_lines) next_line();            // lines behind the item's last unit
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afree + 8 * as);
        }
        if (warp == 0) prof_out(0);
    } else if (warp < 4 + kEpiWarps) {
        // ---------------- epilogue: lane = TMEM lane within the quarter wq = strand row of the tile; the two warps of a
        // quarter take the even and the odd effect sets (a lone warp per quarter recombining ten sets took ~3 300
        // cycles per block, four times what the MMAs of a 270-marker block take)
        const uint32_t wq = warp & 3, half = (warp - 4) >> 2;
        const uint32_t my_sets = (n_sets + 1 - half) / 2;               // sets half, half + 2, ...
        double* const stage = out_stage0 + (size_t) (warp - 4) * (out_bytes_warp / 8);
        uint32_t slot = 0, slot_par = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Tc5Range rg = range_of(item);
            const uint32_t tile = item / n_ranges;
            // four blocks [g0, g0 + 4) staged for the quarter's 32 rows -> 32 contiguous bytes per row and set
            auto flush = [&](uint32_t g_hi) {
                const uint32_t g0 = (g_hi - 1) & ~3u, g_lo = max(g0, rg.blk_lo);
                __syncwarp();
                const uint32_t sl = lane & 3, b = g0 + sl;
                if (b >= g_lo && b < g_hi && !(dbg & 64u)) {
#pragma unroll
                    for (uint32_t i = 0; i < 4; ++i) {
                        const uint32_t rl = 8 * i + (lane >> 2), strand_row = wq * 32 + rl;
                        if (tile * 64 + (strand_row >> 1) < n) {
                            const size_t o = ((size_t) tile * 128 + strand_row) * n_blocks + b;
                            for (uint32_t j = 0; j < my_sets; ++j) P.out[2 * j + half][o] = stage[(j * 32 + rl) * kOutRow + sl];
                        }
                    }
                }
                __syncwarp();
            };
            uint32_t blk_next = rg.blk_lo;
            auto empty_until = [&](uint32_t blk) {                // blocks without markers sum to 0
                while (blk_next < blk) {
                    for (uint32_t j = 0; j < my_sets; ++j) stage[(j * 32 + lane) * kOutRow + (blk_next & 3)] = 0.0;
                    ++blk_next;
                    if ((blk_next & 3) == 0 || blk_next == rg.blk_hi) flush(blk_next);
                }
            };
            // the block numbers two runs ahead and the bases one run ahead: global loads, an L2 round trip each, that a lone
            // warp cannot hide otherwise (they were ~1 400 of the ~3 000 cycles a block's flush took)
            constexpr uint32_t kMySets = (kLocalTc5MaxSets + 1) / 2;
            uint32_t blk_n1 = __ldg(runs + rg.run_lo), blk_n2 = rg.run_lo + 1 < rg.run_hi ? __ldg(runs + rg.run_lo + 1) : 0u;
            double base_n1[kMySets];
#pragma unroll
            for (uint32_t j = 0; j < kMySets; ++j) base_n1[j] = j < my_sets ? __ldg(P.base[2 * j + half] + blk_n1) : 0.0;
#pragma unroll 1
            for (uint32_t run = rg.run_lo; run < rg.run_hi; ++run) {
                const uint32_t blk = blk_n1;
                double base[kMySets];
#pragma unroll
                for (uint32_t j = 0; j < kMySets; ++j) base[j] = base_n1[j];
                blk_n1 = blk_n2;
                if (run + 1 < rg.run_hi) {
#pragma unroll
                    for (uint32_t j = 0; j < kMySets; ++j) if (j < my_sets) base_n1[j] = __ldg(P.base[2 * j + half] + blk_n1);
                }
                if (run + 2 < rg.run_hi) blk_n2 = __ldg(runs + run + 2);
                empty_until(blk);
                wait_on(bar_dready + 8 * slot, slot_par, 0);
                tc_fence_after();
                const uint32_t taddr = tmem + ((wq * 32) << 16) + kSlotCols * slot;
                const uint32_t todo = (dbg & 4u) ? 0u : my_sets;
#pragma unroll
                for (uint32_t j0 = 0; j0 < kMySets + 1; j0 += 3) {
                    // three sets at a time: the loads first, then three independent recombinations
                    uint32_t d[3][8];
                    if (j0 < todo) {
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=r"(d[j][0]), "=r"(d[j][1]), "=r"(d[j][2]), "=r"(d[j][3]), "=r"(d[j][4]), "=r"(d[j][5]), "=r"(d[j][6]), "=r"(d[j][7])
                                         : "r"(taddr + 8 * (2 * min(j0 + j, my_sets - 1) + half)) : "memory");
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    if (j0 + 3 >= todo && (j0 < todo || j0 == 0)) {     // the accumulator can take another block
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_dfree + 8 * slot);
                    }
                    if (j0 < todo) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            if (j0 + j < kMySets) {
                                // sum_k d_k 256^k by Horner in fp64: |d_k| < 2^31, so every step is exact or rounds at 2^-53 relative
                                double v = (double) (int) d[j][7];
#pragma unroll
                                for (int k = 6; k >= 0; --k) v = fma(v, 256.0, (double) (int) d[j][k]);
                                if (j0 + j < my_sets)
                                    stage[((j0 + j) * 32 + lane) * kOutRow + (blk & 3)] = fma(v, P.scale[2 * (j0 + j) + half], base[j0 + j]);
                            }
                        }
                    }
                }
                if (++slot == kSlots) { slot = 0; slot_par ^= 1; }
                blk_next = blk + 1;
                if ((blk_next & 3) == 0 || blk_next == rg.blk_hi) flush(blk_next);
            }
            empty_until(rg.blk_hi);
        }
        if (warp == 4) prof_out(1);
    } else if (warp == kWarpIssuer || warp == kWarpIssuer + 1) {
        // ---------------- issuers: TWO warps take the units in turn.  Each walks its unit's records CONVERGED and one elected
        // lane issues: from a lone thread inside `if (lane == 0)` every tcgen05.mma / commit is wrapped in a waterfall loop
        // (ELECT, R2UR.BROADCAST, BRA.U.ANY) — 47 cycles per MMA and ~300 per commit whatever N; converged + elect.sync the MMA
        // issues in 17 cycles and N = 80 runs at its 40-cycle floor (profiles/r2_umma_i8_probe2.txt).  Why two: the pipe
        // accepts the next MMA only when the one before is nearly done, so a unit's MMAs hold their warp ~50 cycles each, and
        // what that warp does per unit besides (waits for B and the operand stage, records, commits: ~1 100 cycles) left the
        // pipe idle for longer than the MMAs ran.  Now one warp prepares unit u + 1 while the other issues unit u; a
        // barrier hands the pipe over so that the units' MMAs are issued in order (a block's first MMA, which overwrites the
        // accumulator, must precede the rest; completion is in issue order, so a block's commit by one warp covers the
        // other's earlier MMAs too).
        const uint32_t me = warp - kWarpIssuer;
        uint32_t q_unit = 0, ring_i = 0, b_stage = 0, b_par = 0;
        uint32_t q_blk = 0, slot = 0, slot_par = 0;       // blocks ended so far; slot = q_blk % kSlots, slot_par = (q_blk / kSlots) & 1
        // (addresses that came out of shared memory or a cvta, re-declared warp-uniform for the compiler)
        const uint32_t u_tmem = __shfl_sync(0xffffffffu, tmem, 0), u_b_smem = __shfl_sync(0xffffffffu, b_smem, 0);
        const uint32_t d_hi = (uint32_t) (b_descriptor(0) >> 32);
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Tc5Range rg = range_of(item);
            // how many blocks end in a unit this warp skips: from the head of its own unit before (the last word); for an item
            // that starts with the other warp's unit, from the unit table (one exposed L2 round trip per item)
            uint32_t n_ends = ((q_unit & 1u) != me) ? __ldg(&units[rg.unit_lo].n_ends) : 0u;
#pragma unroll 1
            for (uint32_t u = 0; u < rg.n_units; ++u, ++q_unit) {
                if ((q_unit & 1u) != me) {
                    // the other warp's unit: only the accumulator rotation moves on
                    q_blk += n_ends;
                    slot += n_ends % kSlots;
                    slot_par ^= (n_ends / kSlots) & 1u;
                    if (slot >= kSlots) { slot -= kSlots; slot_par ^= 1; }
                } else {
                    wait_on(bar_bfull + 8 * b_stage, b_par, 0);
                    const uint32_t b_addr = u_b_smem + b_stage * stage_bytes;
                    // (records go through a shuffle from lane 0: it tells the compiler they are warp-uniform, and the MMA operands
                    // derived from them then live in uniform registers instead of being moved there one by one)
                    const uint32_t n_seg = __shfl_sync(0xffffffffu, lds32(b_addr), 0);
                    uint32_t d_lo = (uint32_t) b_descriptor(b_addr + kUnitHead);      // the half with the start address
                    uint32_t seg = __shfl_sync(0xffffffffu, lds32(b_addr + 4), 0);
                    n_ends = __shfl_sync(0xffffffffu, lds32(b_addr + kUnitHead - 4), 0);       // of the NEXT unit, which this warp skips
                    wait_on(bar_tfull + 8 * (q_unit & 1), (q_unit >> 1) & 1, 1);
                    tc_fence_after();
                    const uint32_t a_base = u_tmem + kACol0 + kStageCols * (q_unit & 1);
                    uint32_t d_addr = u_tmem + kSlotCols * slot;       // a block that continues from the unit before
                    // the pipe is mine once the other warp has issued the unit before
                    if (q_unit > 0) wait_on(bar_issued + 8 * ((q_unit - 1) & 1), ((q_unit - 1) >> 1) & 1, 2);
                    // a segment = consecutive words of one block, issued as one straight-line burst: a lone warp pays ~15 cycles
                    // per branch, so per-STEP control flow (the first version of this loop) cost 200 cycles per step
#pragma unroll 1
                    for (uint32_t sg = 0; sg < n_seg; ++sg) {
                        const uint32_t seg_next = __shfl_sync(0xffffffffu, lds32(b_addr + 8 + 4 * sg), 0);
                        const uint32_t cnt = (seg >> 4) & 31u;
                        uint32_t acc = 1;
                        if (seg & kSegFirst) {
                            if (q_blk >= kSlots) {                       // the epilogue has read this accumulator's previous sums
                                wait_on(bar_dfree + 8 * slot, slot_par ^ 1, 2);
                                tc_fence_after();
                            }
                            d_addr = u_tmem + kSlotCols * slot;
                            acc = 0;
                        }
                        const uint32_t a_addr = a_base + 8 * (seg & 15u);
                        if (!(dbg & 2u) && elect_one()) {
                            // Straight-line, no predicates: the first MMA of the segment, then a jump into a descending chain of
                            // the remaining cnt - 1 (their order does not matter: they all accumulate)
                            const uint32_t sd = step_bytes >> 4;          // the descriptor's start-address field counts 16-byte units
                            umma_i8_ts_lohi(d_addr, a_addr, d_lo, d_hi, idesc, acc);
                            uint32_t a = a_addr + 8 * (cnt - 1), dl = d_lo + (cnt - 1) * sd;
#define GSB_TC5_STEP(K) case K: umma_i8_ts_lohi(d_addr, a, dl, d_hi, idesc, 1u); a -= 8; dl -= sd;
                            switch (cnt - 1) {
                                GSB_TC5_STEP(19) GSB_TC5_STEP(18) GSB_TC5_STEP(17) GSB_TC5_STEP(16) GSB_TC5_STEP(15) GSB_TC5_STEP(14) GSB_TC5_STEP(13)
                                GSB_TC5_STEP(12) GSB_TC5_STEP(11) GSB_TC5_STEP(10) GSB_TC5_STEP(9) GSB_TC5_STEP(8) GSB_TC5_STEP(7) GSB_TC5_STEP(6)
                                GSB_TC5_STEP(5) GSB_TC5_STEP(4) GSB_TC5_STEP(3) GSB_TC5_STEP(2) GSB_TC5_STEP(1)
                                default: break;
                            }
#undef GSB_TC5_STEP
                        }
                        __syncwarp();
                        d_lo += cnt * (step_bytes >> 4);
                        if (seg & kSegLast) {
                            if (elect_one()) tc_commit(bar_dready + 8 * slot);
                            __syncwarp();
                            ++q_blk;
                            if (++slot == kSlots) { slot = 0; slot_par ^= 1; }
                        }
                        seg = seg_next;
                    }
                    if (lane == 0) mbar_arrive(bar_issued + 8 * (q_unit & 1));       // the pipe is the other warp's
                    if (elect_one()) tc_commit(bar_udone + 8 * ring_i);
                    __syncwarp();
                }
                if (++ring_i == n_ring) ring_i = 0;
                if (++b_stage == n_bstages) { b_stage = 0; b_par ^= 1; }
            }
        }
        if (me == 0) prof_out(2);
    } else if (warp == kWarpALoader) {
        // ---------------- A loader: lanes 8k .. 8k + 7 copy one 128-byte line of one strand row
        uint32_t q_line = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Tc5Range rg = range_of(item);
            const uint32_t tile = item / n_ranges;
            __syncwarp();
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t r = 4 * lane + k, ind = min(tile * 64 + (r >> 1), n - 1);
                s_rowptr[r] = store + (uint64_t) __ldg(rows + ind) * row_words + (r & 1) * plane_words;
            }
            __syncwarp();
#pragma unroll 1
            for (uint32_t line = 0; line < rg.n_lines; ++line, ++q_line) {
                const uint32_t as = q_line % kALines;
                if (q_line >= kALines) wait_on(bar_afree + 8 * as, ((q_line / kALines) - 1) & 1, 0);
                const uint32_t word0 = (rg.line_lo + line) * 32 + (lane & 7) * 4;
                const uint32_t dst0 = a_smem + as * kALineBytes + (lane >> 3) * kARowBytes + (lane & 7) * 16;
#pragma unroll 8
                for (uint32_t j = 0; j < ((dbg & 8u) ? 0u : 32u); ++j) cp_async16_s(dst0 + j * 4 * kARowBytes, s_rowptr[4 * j + (lane >> 3)] + word0);
                cp_async_arrive_noinc(bar_afull + 8 * as);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        prof_out(3);
    } else {
        // ---------------- B loader: one thread, one bulk copy per unit
        if (lane == 0) {
            uint32_t b_stage = 0, b_par = 0, filled = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const Tc5Range rg = range_of(item);
#pragma unroll 1
                for (uint32_t u = 0; u < rg.n_units; ++u) {
                    const uint2 ub = __ldg(reinterpret_cast<const uint2*>(units + rg.unit_lo + u));
                    // stage b_stage was last read by the unit n_bstages back, committed to ring barrier b_stage
                    if (filled >= n_bstages) {
                        const long long t = profiling ? clock64() : 0;
                        mbar_wait(bar_udone + 8 * b_stage, b_par ^ 1, ok);
                        if (profiling) p_wait[0] += clock64() - t;
                    }
                    bulk_g2s(b_smem + b_stage * stage_bytes, bstream + (size_t) ub.x * 16, (dbg & 16u) ? kUnitHead : ub.y, bar_bfull + 8 * b_stage);
                    ++filled;
                    if (++b_stage == n_bstages) { b_stage = 0; b_par ^= 1; }
                }
            }
            prof_out(4);
        }
        __syncwarp();
    }
    if (!ok) atomicExch(flag, 2);

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpIssuer) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// B operand of one effect set: 256 bytes per step.  K index k = 4s + b contracts marker bit 8b + s of the
// step's word (operand byte value 2^s); the byte of (k, digit) sits at (k >> 4) * 128 + digit * 16 + (k & 15).
// Digits are those of round(delta * 2^(exp2 - s)); markers of the word outside the step's block stay zero.
void build_local_tc5_fragments(const std::vector<double>& delta /* [M] */, const std::vector<uint2>& steps,
                               const uint32_t* block_offsets, const uint32_t* block_markers, int exp2, std::vector<uint32_t>& frag) {
    frag.assign(steps.size() * 64, 0u);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(frag.data());
    for (size_t st = 0; st < steps.size(); ++st) {
        const uint32_t w = steps[st].x, blk = steps[st].y & 0x7fffffffu;
        const uint32_t m_lo = block_markers[block_offsets[blk]], m_hi = m_lo + (block_offsets[blk + 1] - block_offsets[blk]);
        for (uint32_t bit = 0; bit < 32; ++bit) {
            const uint32_t m = w * 32 + bit;
            if (m < m_lo || m >= m_hi) continue;
            const uint32_t b = bit >> 3, s = bit & 7, k = 4 * s + b;
            long long q = std::isfinite(delta[m]) ? std::llrint(std::ldexp(delta[m], exp2 - (int) s)) : 0;
            for (uint32_t dg = 0; dg < 8; ++dg) {
                const long long d8 = ((q + 128) & 255) - 128;
                q = (q - d8) >> 8;
                bytes[st * 256 + (k >> 4) * 128 + dg * 16 + (k & 15)] = (uint8_t) (int8_t) d8;
            }
        }
    }
}

// One pass over the genotypes for n_sets (<= kLocalTc5MaxSets) effect sets: d_b / d_base / d_out / scale per
// set; `steps` (marker word, block) in block order and `block_first` (first step of every non-empty block,
// then the number of steps) are the schedule of evaluate.cu.  Every element of the outputs is written.
// Asynchronous.
int local_tc5_launch(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, uint32_t n_blocks,
                     const std::vector<uint2>& steps, const std::vector<uint32_t>& block_first,
                     uint32_t n_sets, const uint32_t* const* d_b, const double* const* d_base, double* const* d_out, const double* scale) {
    GSB_REQUIRE(n_sets >= 1 && n_sets <= kLocalTc5MaxSets, GSB200_ERR_ARG, "local GEBV (tcgen05 path): %u effect sets in one pass", n_sets);
    LocalTc5Params P;
    memset(&P, 0, sizeof P);
    for (uint32_t q = 0; q < n_sets; ++q) { P.b[q] = (const uint4*) d_b[q]; P.base[q] = d_base[q]; P.out[q] = d_out[q]; P.scale[q] = scale[q]; }
    const uint32_t n_steps = (uint32_t) steps.size(), n_runs = (uint32_t) block_first.size() - 1;
    const uint32_t n_tiles = (n + 63) / 64;
    const uint32_t step_bytes = 256 * n_sets;

    // ranges of whole blocks: enough items for ~8 per SM (few tiles) but at least 64 steps each
    uint32_t want = (uint32_t) std::max<uint64_t>(1, ((uint64_t) ctx->sm_count * 8 + n_tiles - 1) / n_tiles);
    want = std::min(want, std::min(n_runs, std::max(1u, n_steps / 64)));
    std::vector<uint32_t> cut(1, 0u);                       // first run of every range
    for (uint32_t r = 1; r < want; ++r) {
        const uint32_t target = (uint32_t) ((uint64_t) n_steps * r / want);
        const uint32_t run = (uint32_t) (std::lower_bound(block_first.begin(), block_first.end(), target) - block_first.begin());
        if (run > cut.back() && run < n_runs) cut.push_back(run);
    }
    cut.push_back(n_runs);
    const uint32_t n_ranges = (uint32_t) cut.size() - 1;
    std::vector<Tc5Range> ranges(n_ranges);
    std::vector<Tc5Unit> units;
    std::vector<uint32_t> runs(n_runs), unit_step, heads;
    for (uint32_t r = 0; r < n_runs; ++r) runs[r] = steps[block_first[r]].y & 0x7fffffffu;
    uint64_t b_off16 = 0;
    for (uint32_t k = 0; k < n_ranges; ++k) {
        const uint32_t s_lo = block_first[cut[k]], s_hi = block_first[cut[k + 1]];
        Tc5Range& g = ranges[k];
        g.unit_lo = (uint32_t) units.size();
        g.run_lo = cut[k]; g.run_hi = cut[k + 1];
        g.line_lo = steps[s_lo].x >> 5; g.n_lines = (steps[s_hi - 1].x >> 5) - g.line_lo + 1;
        g.blk_lo = k == 0 ? 0u : runs[cut[k]];
        g.blk_hi = k + 1 == n_ranges ? n_blocks : runs[cut[k + 1]];
        uint32_t line_at = g.line_lo - 1;                   // the line the expanders hold (none yet: one before the first)
        for (uint32_t s = s_lo; s < s_hi;) {
            // a unit: the steps of one 16-word half line, at most kUnitSteps of them
            const uint32_t half = steps[s].x / kUnitWords;
            uint32_t e = s;
            while (e < s_hi && e - s < kUnitSteps && steps[e].x / kUnitWords == half) ++e;
            Tc5Unit u;
            u.b_off16 = (uint32_t) b_off16;
            u.b_bytes = kUnitHead + (e - s) * step_bytes;
            u.line_adv_half = (((half >> 1) - line_at) << 1) | (half & 1u);
            u.n_ends = 0;
            line_at = half >> 1;
            b_off16 += u.b_bytes / 16;
            GSB_REQUIRE(b_off16 < (1ull << 32), GSB200_ERR_ARG, "local GEBV (tcgen05 path): B stream too long");
            units.push_back(u);
            unit_step.push_back(s);
            const size_t h0 = heads.size();
            heads.resize(h0 + kUnitHead / 4, 0u);
            uint32_t n_seg = 0;
            for (uint32_t p = s; p < e;) {
                const uint32_t blk = steps[p].y & 0x7fffffffu;
                uint32_t r = p + 1;
                while (r < e && (steps[r].y & 0x7fffffffu) == blk) ++r;
                const bool first = p == s_lo || (steps[p - 1].y & 0x7fffffffu) != blk;
                const bool last = r == s_hi || (steps[r].y & 0x7fffffffu) != blk;
                heads[h0 + 1 + n_seg++] = (steps[p].x % kUnitWords) | ((r - p) << 4) | (first ? kSegFirst : 0u) | (last ? kSegLast : 0u);
                units.back().n_ends += last;
                p = r;
            }
            heads[h0] = n_seg;
            s = e;
        }
        g.n_units = (uint32_t) units.size() - g.unit_lo;
        for (uint32_t q = g.unit_lo; q + 1 < (uint32_t) units.size(); ++q) heads[(size_t) q * (kUnitHead / 4) + kUnitHead / 4 - 1] = units[q + 1].n_ends;
    }
    unit_step.push_back(n_steps);
    const uint32_t n_units = (uint32_t) units.size();

    // plan -> device: ranges | units | runs | first step of every unit | heads, then the B stream
    LocalSchedule& sc = ctx->local_sched;
    auto pad16 = [](size_t x) { return (x + 15) & ~(size_t) 15; };
    const size_t off_units = ranges.size() * sizeof(Tc5Range), off_runs = off_units + units.size() * sizeof(Tc5Unit);
    const size_t off_us = pad16(off_runs + runs.size() * 4), off_heads = pad16(off_us + unit_step.size() * 4);
    const size_t plan_bytes = off_heads + heads.size() * 4;
    std::vector<uint8_t> plan(plan_bytes);
    memcpy(plan.data(), ranges.data(), off_units);
    memcpy(plan.data() + off_units, units.data(), units.size() * sizeof(Tc5Unit));
    memcpy(plan.data() + off_runs, runs.data(), runs.size() * 4);
    memcpy(plan.data() + off_us, unit_step.data(), unit_step.size() * 4);
    memcpy(plan.data() + off_heads, heads.data(), heads.size() * 4);
    GSB_TRY(sc.d_plan.ensure(plan_bytes));
    GSB_TRY(sc.d_bstream.ensure((size_t) b_off16 * 16));
    GSB_CUDA_TRY(cudaMemcpyAsync(sc.d_plan.ptr, plan.data(), plan_bytes, cudaMemcpyHostToDevice, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));          // `plan` is pageable host memory about to go out of scope
    const char* dp = (const char*) sc.d_plan.ptr;
    const Tc5Unit* d_units = (const Tc5Unit*) (dp + off_units);
    local_tc5_pack_kernel<<<n_units, 256, 0, ctx->stream>>>((uint4*) sc.d_bstream.ptr, d_units, n_sets,
                                                             (const uint32_t*) (dp + off_us), (const uint32_t*) (dp + off_heads), P);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());

    const uint64_t n_items64 = (uint64_t) n_tiles * n_ranges;
    GSB_REQUIRE(n_items64 < (1ull << 31), GSB200_ERR_ARG, "local GEBV: too many work items");
    const uint32_t n_items = (uint32_t) n_items64;
    const uint32_t fixed = kALines * kALineBytes + kEpiWarps * ((n_sets + 1) / 2) * 32 * kOutRow * 8;
    const uint32_t stage_bytes = kUnitHead + kUnitSteps * step_bytes + kStageSlack;
    const uint32_t n_bstages = std::min<uint32_t>(kMaxBStages, (kSmemBudget - fixed) / stage_bytes);
    GSB_REQUIRE(n_bstages >= 2, GSB200_ERR_ARG, "local GEBV (tcgen05 path): internal: shared memory does not hold two B stages");
    const uint32_t smem = fixed + n_bstages * stage_bytes;
    GSB_CUDA_TRY(cudaFuncSetAttribute(local_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBudget));
    static const uint32_t dbg = getenv("GSB200_LOCAL_DEBUG") ? (uint32_t) atoi(getenv("GSB200_LOCAL_DEBUG")) : 0u;
    const uint32_t grid = std::min<uint32_t>((uint32_t) ctx->sm_count, n_items);
    // GSB200_LOCAL_PROF=<file>: where CTA 0's roles spent their cycles (one line per launch)
    static const char* prof_path = getenv("GSB200_LOCAL_PROF");
    long long* d_prof = nullptr;
    if (prof_path) {
        GSB_CUDA_TRY(cudaMalloc((void**) &d_prof, 20 * sizeof(long long)));
        GSB_CUDA_TRY(cudaMemsetAsync(d_prof, 0, 20 * sizeof(long long), ctx->stream));
    }
    local_tc5_kernel<<<grid, kThreads, smem, ctx->stream>>>(ctx->d_store, ctx->row_words(), ctx->plane_words, d_rows, n, n_blocks,
                                                           (const uint8_t*) sc.d_bstream.ptr, (const Tc5Range*) dp, n_ranges, d_units, (const uint32_t*) (dp + off_runs),
                                                           n_items, n_sets, n_bstages, P, ctx->d_flag + 2, dbg, d_prof);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    if (d_prof) {
        long long h[20];
        GSB_CUDA_TRY(cudaMemcpyAsync(h, d_prof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_prof);
        if (FILE* f = fopen(prof_path, "a")) {
            fprintf(f, "items %u (tiles %u x ranges %u) steps %u units %u sets %u B stages %u | cycles of CTA 0: "
                       "expander total %lld wait(line) %lld wait(stage free) %lld | epilogue total %lld wait(sums) %lld | "
                       "issuer total %lld wait(B) %lld wait(operand) %lld wait(accumulator free) %lld | A loader total %lld wait(line free) %lld | "
                       "B loader total %lld wait(stage free) %lld\n", n_items, n_tiles, n_ranges, n_steps, n_units, n_sets, n_bstages,
                    h[0], h[1], h[2], h[4], h[5], h[8], h[9], h[10], h[11], h[12], h[13], h[16], h[17]);
            fclose(f);
        }
    }
    return GSB200_OK;
}

}  // namespace gsb
