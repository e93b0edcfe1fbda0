// store.cu — the device-resident bit-packed haplotype store and its chars<->bits boundary (K8).
//
// Replaces the payload of gsc_AlleleMatrix (core.h:776-806): instead of one malloc'ed
// char[2*M] per genotype, one row of 2 haplotypes x `planes` bit-planes, 128-byte aligned.
// Meiosis only ever copies bits, so it is symbol-agnostic; symbols matter only here (pack /
// unpack) and in the evaluation kernels, through the per-marker dictionary.
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cstring>

using namespace gsb;

namespace {

// rows[i] <- chars[i]: one thread per (row, 32-marker word)
__global__ void pack_kernel(uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
                            uint32_t planes, uint32_t n_markers,
                            const uint32_t* __restrict__ rows, uint32_t n,
                            const uint8_t* __restrict__ chars,
                            const uint8_t* __restrict__ dict_sym, const uint16_t* __restrict__ dict_n,
                            int* __restrict__ unknown /* raised when a symbol is not in its marker's dictionary */) {
    const uint32_t words = (n_markers + 31) / 32;
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) n * words) return;
    const uint32_t i = (uint32_t) (tid / words), w = (uint32_t) (tid % words);
    const uint8_t* src = chars + (size_t) i * 2 * n_markers;
    uint32_t acc[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[h][p] = 0;
    const uint32_t m_end = min(n_markers, (w + 1) * 32);
    for (uint32_t m = w * 32; m < m_end; ++m) {
        const uint8_t* sym = dict_sym + ((size_t) m << planes);
        const uint32_t nc = dict_n[m];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint8_t a = src[2 * (size_t) m + h];
            uint32_t code = 0xffffffffu;
            for (uint32_t c = 0; c < nc; ++c) if (sym[c] == a) { code = c; break; }
            if (code == 0xffffffffu) { *unknown = 1; code = 0; }
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[h][p] |= ((code >> p) & 1u) << (m & 31);
        }
    }
    uint32_t* dst = store + (uint64_t) rows[i] * row_words;
    for (uint32_t h = 0; h < 2; ++h)
        for (uint32_t p = 0; p < planes; ++p)
            dst[(uint64_t) (h * planes + p) * plane_words + w] = acc[h][p];
}

// chars[i] <- markers [m_lo, m_hi) of rows[i]; the output row is 2 * (m_hi - m_lo) chars
__global__ void unpack_kernel(const uint32_t* __restrict__ store, uint64_t row_words, uint32_t plane_words,
                              uint32_t planes, uint32_t m_lo, uint32_t m_hi,
                              const uint32_t* __restrict__ rows, uint32_t n,
                              uint8_t* __restrict__ chars,
                              const uint8_t* __restrict__ dict_sym) {
    const uint32_t w_lo = m_lo >> 5, words = ((m_hi + 31) >> 5) - w_lo;
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) n * words) return;
    const uint32_t i = (uint32_t) (tid / words), w = w_lo + (uint32_t) (tid % words);
    const uint32_t* src = store + (uint64_t) rows[i] * row_words;
    uint32_t bits[2][8];
    for (uint32_t h = 0; h < 2; ++h)
        for (uint32_t p = 0; p < 8; ++p)
            bits[h][p] = p < planes ? src[(uint64_t) (h * planes + p) * plane_words + w] : 0u;
    uint8_t* dst = chars + (size_t) i * 2 * (m_hi - m_lo);
    const uint32_t m_end = min(m_hi, (w + 1) * 32);
    for (uint32_t m = max(w * 32, m_lo); m < m_end; ++m) {
        const uint8_t* sym = dict_sym + ((size_t) m << planes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t code = 0;
#pragma unroll
            for (int p = 0; p < 8; ++p) code |= ((bits[h][p] >> (m & 31)) & 1u) << p;
            dst[2 * (size_t) (m - m_lo) + h] = sym[code];
        }
    }
}

// copy every row into a store with more planes (new high planes are zero)
__global__ void widen_kernel(const uint32_t* __restrict__ old_store, uint32_t* __restrict__ new_store,
                             uint32_t plane_words, uint32_t old_planes, uint32_t new_planes, uint64_t n_rows) {
    const uint64_t per_row = (uint64_t) 2 * new_planes * plane_words;
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_rows * per_row) return;
    const uint64_t row = tid / per_row, r = tid % per_row;
    const uint32_t h = (uint32_t) (r / ((uint64_t) new_planes * plane_words));
    const uint32_t p = (uint32_t) ((r / plane_words) % new_planes);
    const uint32_t w = (uint32_t) (r % plane_words);
    uint32_t v = 0;
    if (p < old_planes) v = old_store[row * 2 * old_planes * plane_words + (uint64_t) (h * old_planes + p) * plane_words + w];
    new_store[tid] = v;
}

// packed rows <-> a dense device buffer (n x row_words), e.g. an NCCL send/receive buffer
__global__ void move_rows_kernel(uint32_t* store, uint32_t* dense, const uint32_t* rows, uint32_t n, uint64_t row_words, int to_store) {
    const uint64_t per_row = row_words / 4;
    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (uint64_t) n * per_row) return;
    const uint32_t i = (uint32_t) (tid / per_row);
    const uint64_t w = tid % per_row;
    uint4* a = reinterpret_cast<uint4*>(store + (uint64_t) rows[i] * row_words) + w;
    uint4* b = reinterpret_cast<uint4*>(dense + (uint64_t) i * row_words) + w;
    if (to_store) *a = *b; else *b = *a;
}

int move_rows(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, void* dense, int to_store, const char* what) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(dense != nullptr && d_rows != nullptr, GSB200_ERR_ARG, "%s: null device buffer", what);
    GSB_REQUIRE(((uintptr_t) dense & 15) == 0, GSB200_ERR_ARG, "%s: device buffer must be 16-byte aligned", what);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t threads = (uint64_t) n * (ctx->row_words() / 4);
    move_rows_kernel<<<(unsigned) ((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_store, (uint32_t*) dense, d_rows, n,
                                                                                  ctx->row_words(), to_store);
    ++ctx->launches;
    GSB_CUDA_TRY(cudaGetLastError());
    return GSB200_OK;
}

int widen_store(gsb200_ctx* ctx, uint32_t new_planes) {
    const uint32_t old_planes = ctx->planes;
    // host dictionary
    std::vector<uint8_t> nd((size_t) ctx->n_markers << new_planes, 0);
    for (uint32_t m = 0; m < ctx->n_markers; ++m)
        memcpy(&nd[(size_t) m << new_planes], &ctx->dict_sym[(size_t) m << old_planes], (size_t) 1 << old_planes);
    ctx->dict_sym.swap(nd);
    ctx->dict_dirty = true;
    // device rows
    if (ctx->capacity > 0) {
        uint32_t* fresh = nullptr;
        const uint64_t words = ctx->capacity * 2 * new_planes * ctx->plane_words;
        GSB_CUDA_TRY(cudaMalloc((void**) &fresh, words * sizeof(uint32_t)));
        const uint32_t threads = 256;
        const uint64_t blocks = (words + threads - 1) / threads;
        widen_kernel<<<(unsigned) blocks, threads, 0, ctx->stream>>>(ctx->d_store, fresh, ctx->plane_words,
                                                                      old_planes, new_planes, ctx->capacity);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_store);
        ctx->d_store = fresh;
    }
    ctx->planes = new_planes;
    ++ctx->dict_version;
    return GSB200_OK;
}

}  // namespace

namespace gsb {

// code of `symbol` at `marker`, added to the dictionary (widening the store if its codes are
// used up) when it is not there yet
int ensure_symbol(gsb200_ctx* ctx, uint32_t marker, uint8_t symbol, uint32_t* code) {
    const uint32_t nc = ctx->dict_n[marker];
    const uint8_t* sym = &ctx->dict_sym[(size_t) marker << ctx->planes];
    for (uint32_t c = 0; c < nc; ++c) if (sym[c] == symbol) { *code = c; return GSB200_OK; }
    if (nc == (1u << ctx->planes)) {
        GSB_REQUIRE(ctx->planes < 8, GSB200_ERR_ARG, "store: more than 256 symbols at marker %u", marker);
        GSB_TRY(widen_store(ctx, ctx->planes * 2));
    }
    ctx->dict_sym[((size_t) marker << ctx->planes) + nc] = symbol;
    ctx->dict_n[marker] = (uint16_t) (nc + 1);
    ctx->dict_dirty = true;
    ++ctx->dict_version;
    *code = nc;
    return GSB200_OK;
}

int sync_dictionary(gsb200_ctx* ctx) {
    if (!ctx->dict_dirty && ctx->d_dict_sym) return GSB200_OK;
    const size_t bytes = (size_t) ctx->n_markers << ctx->planes;
    if (ctx->d_dict_sym) { cudaFree(ctx->d_dict_sym); ctx->d_dict_sym = nullptr; }
    GSB_CUDA_TRY(cudaMalloc((void**) &ctx->d_dict_sym, bytes));
    if (!ctx->d_dict_n) GSB_CUDA_TRY(cudaMalloc((void**) &ctx->d_dict_n, ctx->n_markers * sizeof(uint16_t)));
    GSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_dict_sym, ctx->dict_sym.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    GSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_dict_n, ctx->dict_n.data(), ctx->n_markers * sizeof(uint16_t),
                                 cudaMemcpyHostToDevice, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->dict_dirty = false;
    return GSB200_OK;
}

}  // namespace gsb

extern "C" {

int gsb200_store_reserve(gsb200_ctx* ctx, uint64_t n_rows) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n_rows <= ctx->capacity) return GSB200_OK;
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    // grow geometrically so that repeated small appends stay amortised O(1)
    uint64_t want = std::max<uint64_t>(n_rows, ctx->capacity + ctx->capacity / 2);
    want = std::max<uint64_t>(want, 64);
    const uint64_t row_bytes = ctx->row_words() * sizeof(uint32_t);
    uint32_t* fresh = nullptr;
    cudaError_t e = cudaMalloc((void**) &fresh, want * row_bytes);
    if (e != cudaSuccess && want > n_rows) {   // retry with the exact request
        cudaGetLastError();
        want = n_rows;
        e = cudaMalloc((void**) &fresh, want * row_bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("store: cannot allocate %llu rows of %llu bytes: %s", (unsigned long long) want,
                  (unsigned long long) row_bytes, cudaGetErrorString(e));
        return GSB200_ERR_OOM;
    }
    if (ctx->capacity > 0) {
        GSB_CUDA_TRY(cudaMemcpyAsync(fresh, ctx->d_store, ctx->capacity * row_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // rows start zeroed, as the reference's slots do (core.c:75)
    GSB_CUDA_TRY(cudaMemsetAsync((char*) fresh + ctx->capacity * row_bytes, 0, (want - ctx->capacity) * row_bytes, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_store) cudaFree(ctx->d_store);
    ctx->d_store = fresh;
    ctx->capacity = want;
    return GSB200_OK;
}

uint64_t gsb200_store_capacity(const gsb200_ctx* ctx) { return ctx ? ctx->capacity : 0; }
uint32_t gsb200_store_planes(const gsb200_ctx* ctx) { return ctx ? ctx->planes : 0; }
uint32_t gsb200_store_plane_words(const gsb200_ctx* ctx) { return ctx ? ctx->plane_words : 0; }
uint64_t gsb200_store_row_bytes(const gsb200_ctx* ctx) { return ctx ? ctx->row_words() * sizeof(uint32_t) : 0; }
void* gsb200_store_device_ptr(gsb200_ctx* ctx) { return ctx ? (void*) ctx->d_store : nullptr; }

// Symbols of rows [lo, hi) that their marker's dictionary does not hold yet are added in order of first
// appearance (row by row, marker by marker, strand 0 before strand 1), widening the store when a marker runs
// out of codes.  Rows whose symbols are all known are cleared by a vectorisable pass against the two commonest
// symbols per marker; only rows that fail it take the slow scan.
static int extend_dictionary(gsb200_ctx* ctx, uint32_t lo, uint32_t hi, const char* chars, bool* changed_out) {
    const uint32_t M = ctx->n_markers;
    uint32_t planes = ctx->planes;
    bool changed = false;
    std::vector<uint8_t> e0(2 * (size_t) M), e1(2 * (size_t) M);
    auto refresh_expanded = [&]() {
        for (uint32_t m = 0; m < M; ++m) {
            const uint8_t* sym = &ctx->dict_sym[(size_t) m << ctx->planes];
            const uint32_t nc = ctx->dict_n[m];
            const uint8_t a0 = nc > 0 ? sym[0] : 0, a1 = nc > 1 ? sym[1] : a0;
            e0[2 * (size_t) m] = e0[2 * (size_t) m + 1] = a0;
            e1[2 * (size_t) m] = e1[2 * (size_t) m + 1] = a1;
        }
    };
    refresh_expanded();
    bool any_empty = false;
    for (uint32_t m = 0; m < M && !any_empty; ++m) any_empty = ctx->dict_n[m] == 0;
    for (uint32_t i = lo; i < hi; ++i) {
        const uint8_t* g = (const uint8_t*) chars + (size_t) i * 2 * M;
        if (!any_empty) {
            uint8_t bad = 0;
            const size_t len = 2 * (size_t) M;
            for (size_t j = 0; j < len; ++j) bad |= (uint8_t) ((g[j] != e0[j]) & (g[j] != e1[j]));
            if (!bad) continue;
        }
        bool row_changed = false;
        for (uint32_t m = 0; m < M; ++m) {
            for (uint32_t h = 0; h < 2; ++h) {
                const uint8_t a = g[2 * (size_t) m + h];
                const uint8_t* sym = &ctx->dict_sym[(size_t) m << planes];
                const uint32_t nc = ctx->dict_n[m];
                bool found = false;
                for (uint32_t c = 0; c < nc; ++c) if (sym[c] == a) { found = true; break; }
                if (found) continue;
                if (nc == (1u << planes)) {
                    GSB_REQUIRE(planes < 8, GSB200_ERR_ARG, "store: more than 256 symbols at marker %u", m);
                    GSB_TRY(widen_store(ctx, planes * 2));
                    planes = ctx->planes;
                }
                ctx->dict_sym[((size_t) m << planes) + nc] = a;
                ctx->dict_n[m] = (uint16_t) (nc + 1);
                row_changed = true;
            }
        }
        if (row_changed) {
            changed = true;
            refresh_expanded();
            any_empty = false;
            for (uint32_t m = 0; m < M && !any_empty; ++m) any_empty = ctx->dict_n[m] == 0;
        }
    }
    if (changed) { ctx->dict_dirty = true; ++ctx->dict_version; }
    *changed_out = changed;
    return GSB200_OK;
}

// chars of rows[0 .. n) -> the packed store, in bounded chunks staged on the device; *unknown (host) tells
// whether the pack kernel met a symbol its dictionary does not hold
static int pack_rows(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, const char* chars, int* unknown) {
    const uint32_t M = ctx->n_markers;
    const size_t row_chars = 2 * (size_t) M;
    const uint32_t chunk = (uint32_t) std::max<size_t>(1, std::min<size_t>(n, ((size_t) 256 << 20) / row_chars));
    const uint32_t words = (M + 31) / 32;
    GSB_TRY(sync_dictionary(ctx));
    GSB_CUDA_TRY(cudaMemsetAsync(ctx->d_flag + 3, 0, sizeof(int), ctx->stream));
    uint32_t* d_rows = nullptr;
    GSB_TRY(upload_u32(ctx, ctx->s_rows[0], rows, n, &d_rows));
    GSB_TRY(ctx->s_io.ensure((size_t) chunk * row_chars));
    for (uint32_t done = 0; done < n; done += chunk) {
        const uint32_t cnt = std::min(chunk, n - done);
        // stream-ordered: the copy of chunk k + 1 waits for the pack of chunk k by itself
        GSB_CUDA_TRY(cudaMemcpyAsync(ctx->s_io.ptr, chars + (size_t) done * row_chars, cnt * row_chars, cudaMemcpyHostToDevice, ctx->stream));
        const uint64_t threads = (uint64_t) cnt * words;
        pack_kernel<<<(unsigned) ((threads + 127) / 128), 128, 0, ctx->stream>>>(
            ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes, M, d_rows + done, cnt,
            (const uint8_t*) ctx->s_io.ptr, ctx->d_dict_sym, ctx->d_dict_n, ctx->d_flag + 3);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
    }
    GSB_CUDA_TRY(cudaMemcpyAsync(unknown, ctx->d_flag + 3, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

int gsb200_store_upload_chars(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, const char* chars) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0) return GSB200_OK;
    GSB_REQUIRE(chars != nullptr, GSB200_ERR_ARG, "gsb200_store_upload_chars: null chars");
    GSB_TRY(check_rows(ctx, n, rows, "gsb200_store_upload_chars"));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    // The dictionary is seeded from the first few rows on the host; everything else is packed OPTIMISTICALLY: the
    // pack kernel looks every symbol up on the device and raises a flag for one it does not know, and only then
    // (a new symbol deep inside a batch: rare after the founders' first rows) the host scans the whole batch, extends
    // the dictionary in order of first appearance — the codes come out exactly as a row-by-row scan gives them —
    // and the batch is packed again.  The common case crosses the chars once, at the speed of the PCIe copy.
    bool changed = false;
    GSB_TRY(extend_dictionary(ctx, 0, std::min<uint32_t>(n, 8), chars, &changed));
    int unknown = 0;
    GSB_TRY(pack_rows(ctx, n, rows, chars, &unknown));
    if (unknown) {
        GSB_TRY(extend_dictionary(ctx, 0, n, chars, &changed));
        GSB_TRY(pack_rows(ctx, n, rows, chars, &unknown));
        GSB_REQUIRE(!unknown, GSB200_ERR_ARG, "gsb200_store_upload_chars: internal: a symbol is still missing from the dictionary");
    }
    return GSB200_OK;
}

int gsb200_store_download_chars_range(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, uint32_t first_marker, uint32_t n_markers_out, char* chars) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    if (n == 0 || n_markers_out == 0) return GSB200_OK;
    const char* what = "gsb200_store_download_chars";
    GSB_REQUIRE(chars != nullptr, GSB200_ERR_ARG, "%s: null output", what);
    GSB_REQUIRE((uint64_t) first_marker + n_markers_out <= ctx->n_markers, GSB200_ERR_ARG, "%s: markers %u..%llu are beyond the genome's %u",
                what, first_marker, (unsigned long long) first_marker + n_markers_out, ctx->n_markers);
    GSB_TRY(check_rows(ctx, n, rows, what));
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    GSB_TRY(sync_dictionary(ctx));
    const uint32_t m_lo = first_marker, m_hi = first_marker + n_markers_out;
    const size_t row_chars = 2 * (size_t) n_markers_out;
    const uint32_t chunk = (uint32_t) std::max<size_t>(1, std::min<size_t>(n, ((size_t) 256 << 20) / row_chars));
    const uint32_t words = ((m_hi + 31) >> 5) - (m_lo >> 5);
    for (uint32_t done = 0; done < n; done += chunk) {
        const uint32_t cnt = std::min(chunk, n - done);
        uint32_t* d_rows = nullptr;
        GSB_TRY(upload_u32(ctx, ctx->s_rows[0], rows + done, cnt, &d_rows));
        GSB_TRY(ctx->s_io.ensure(cnt * row_chars));
        const uint64_t threads = (uint64_t) cnt * words;
        unpack_kernel<<<(unsigned) ((threads + 127) / 128), 128, 0, ctx->stream>>>(
            ctx->d_store, ctx->row_words(), ctx->plane_words, ctx->planes, m_lo, m_hi, d_rows, cnt,
            (uint8_t*) ctx->s_io.ptr, ctx->d_dict_sym);
        ++ctx->launches;
        GSB_CUDA_TRY(cudaGetLastError());
        GSB_CUDA_TRY(cudaMemcpyAsync(chars + (size_t) done * row_chars, ctx->s_io.ptr, cnt * row_chars,
                                     cudaMemcpyDeviceToHost, ctx->stream));
        GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return GSB200_OK;
}

int gsb200_store_download_chars(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, char* chars) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    return gsb200_store_download_chars_range(ctx, n, rows, 0, ctx->n_markers, chars);
}

int gsb200_store_gather_rows_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, void* d_dense) {
    return move_rows(ctx, n, d_rows, d_dense, 0, "gsb200_store_gather_rows_dev");
}

int gsb200_store_scatter_rows_dev(gsb200_ctx* ctx, uint32_t n, const uint32_t* d_rows, const void* d_dense) {
    return move_rows(ctx, n, d_rows, (void*) d_dense, 1, "gsb200_store_scatter_rows_dev");
}

int gsb200_store_change_symbol(gsb200_ctx* ctx, uint32_t marker, char from, char to) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_REQUIRE(marker == UINT32_MAX || marker < ctx->n_markers, GSB200_ERR_ARG,
                "gsb200_store_change_symbol: marker %u out of range", marker);
    // Only the dictionary changes: two codes may now decode to the same symbol, which the
    // evaluation tables handle (they are keyed by symbol, not by code).
    const uint32_t lo = marker == UINT32_MAX ? 0 : marker, hi = marker == UINT32_MAX ? ctx->n_markers : marker + 1;
    bool changed = false;
    for (uint32_t m = lo; m < hi; ++m) {
        uint8_t* sym = &ctx->dict_sym[(size_t) m << ctx->planes];
        for (uint32_t c = 0; c < ctx->dict_n[m]; ++c)
            if (sym[c] == (uint8_t) from) { sym[c] = (uint8_t) to; changed = true; }
    }
    if (changed) { ctx->dict_dirty = true; ++ctx->dict_version; }
    return GSB200_OK;
}

int gsb200_store_dictionary(const gsb200_ctx* ctx, char* symbols, uint16_t* n_codes) {
    GSB_REQUIRE(ctx != nullptr && symbols != nullptr && n_codes != nullptr, GSB200_ERR_ARG, "gsb200_store_dictionary: null argument");
    memcpy(symbols, ctx->dict_sym.data(), ctx->dict_sym.size());
    memcpy(n_codes, ctx->dict_n.data(), ctx->dict_n.size() * sizeof(uint16_t));
    return GSB200_OK;
}

int gsb200_store_symbols(const gsb200_ctx* ctx, uint32_t marker, char* out, uint32_t cap) {
    if (!ctx || marker >= ctx->n_markers) return 0;
    const uint32_t nc = ctx->dict_n[marker];
    for (uint32_t c = 0; c < nc && c < cap; ++c) out[c] = (char) ctx->dict_sym[((size_t) marker << ctx->planes) + c];
    return (int) nc;
}

}  // extern "C"
