// context.cu — context lifetime, error reporting, scratch buffers, map and effect-set upload.
#include "gsb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <limits>
#include <numeric>

namespace gsb {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

int Scratch::ensure(size_t need) {
    if (need <= bytes) return GSB200_OK;
    release();
    size_t want = need + need / 4 + 256;
    cudaError_t e = stream ? cudaMallocAsync(&ptr, want, stream) : cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? GSB200_ERR_OOM : GSB200_ERR_CUDA;
    }
    ordered = stream != nullptr;
    bytes = want;
    return GSB200_OK;
}

void Scratch::release() {
    if (ptr) { if (ordered) cudaFreeAsync(ptr, stream); else cudaFree(ptr); }
    ptr = nullptr;
    bytes = 0;
}

void DeviceMap::release() {
    void* ptrs[] = { d_chr, d_dists, d_marker_index, d_bucket, d_chunk_chr, d_chunk_start, d_uncovered };
    for (void* p : ptrs) if (p) cudaFree(p);
    *this = DeviceMap();
}

void DeviceEffects::release() {
    void* ptrs[] = { d_value_all, d_value_first, d_centre, d_delta, d_digit_frag, d_digit_tc5, d_local, d_local_tc5, d_bvq, d_nonfinite };
    for (void* p : ptrs) if (p) cudaFree(p);
    *this = DeviceEffects();
}

int upload_u32(gsb200_ctx* ctx, Scratch& s, const uint32_t* host, size_t n, uint32_t** out) {
    GSB_TRY(s.ensure(std::max<size_t>(n, 1) * sizeof(uint32_t)));
    if (n) GSB_CUDA_TRY(cudaMemcpyAsync(s.ptr, host, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    *out = (uint32_t*) s.ptr;
    return GSB200_OK;
}

int check_rows(gsb200_ctx* ctx, uint32_t n, const uint32_t* rows, const char* what) {
    GSB_REQUIRE(n == 0 || rows != nullptr, GSB200_ERR_ARG, "%s: null row array", what);
    uint32_t top = 0;
    for (uint32_t i = 0; i < n; ++i) top = rows[i] > top ? rows[i] : top;       // vectorises; the loop below only runs to word the message
    if (n == 0 || top < ctx->capacity) return GSB200_OK;
    for (uint32_t i = 0; i < n; ++i) {
        GSB_REQUIRE(rows[i] < ctx->capacity, GSB200_ERR_ARG, "%s: row %u (entry %u) is beyond the store capacity %llu",
                    what, rows[i], i, (unsigned long long) ctx->capacity);
    }
    return GSB200_OK;
}

template <typename T>
static int to_device(const std::vector<T>& v, T** out) {
    *out = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    GSB_CUDA_TRY(cudaMalloc((void**) out, bytes));
    if (!v.empty()) GSB_CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GSB200_OK;
}

}  // namespace gsb

using namespace gsb;

extern "C" {

int gsb200_abi_version(void) { return GSB200_ABI_VERSION; }
const char* gsb200_last_error(void) { return gsb::g_error; }

int gsb200_create(gsb200_ctx** out, int device_ordinal, uint32_t n_markers) {
    GSB_REQUIRE(out != nullptr, GSB200_ERR_ARG, "gsb200_create: null output pointer");
    *out = nullptr;
    GSB_REQUIRE(n_markers > 0, GSB200_ERR_ARG, "gsb200_create: n_markers must be positive");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no usable CUDA device (%s); this engine has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return GSB200_ERR_CUDA;
    }
    GSB_REQUIRE(device_ordinal >= 0 && device_ordinal < count, GSB200_ERR_ARG,
                "gsb200_create: device %d out of range (%d devices)", device_ordinal, count);
    GSB_CUDA_TRY(cudaSetDevice(device_ordinal));
    gsb200_ctx* ctx = new gsb200_ctx();
    ctx->device = device_ordinal;
    ctx->n_markers = n_markers;
    ctx->planes = 1;
    uint32_t words = (n_markers + 31) / 32;
    ctx->plane_words = (words + kPlaneAlignWords - 1) / kPlaneAlignWords * kPlaneAlignWords;
    ctx->dict_sym.assign((size_t) n_markers << ctx->planes, 0);
    ctx->dict_n.assign(n_markers, 0);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) == cudaSuccess) {
        ctx->sm_count = prop.multiProcessorCount;
        ctx->l2_persist_max = (size_t) prop.persistingL2CacheMaxSize;
        ctx->l2_window_max = (size_t) prop.accessPolicyMaxWindowSize;
    }
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete ctx;
        return GSB200_ERR_CUDA;
    }
    // every scratch buffer lives in the stream-ordered pool of this stream; freed blocks stay cached in the pool
    for (auto& sc : ctx->s_rows) sc.stream = ctx->stream;
    for (auto& sc : ctx->s_plan) sc.stream = ctx->stream;
    for (auto& sc : ctx->s_eval) sc.stream = ctx->stream;
    for (auto& sc : ctx->s_chain) sc.stream = ctx->stream;
    for (auto& sc : ctx->s_bv) sc.stream = ctx->stream;
    ctx->s_io.stream = ctx->s_null.stream = ctx->s_select.stream = ctx->local_sched.d_steps.stream = ctx->local_sched.d_plan.stream = ctx->local_sched.d_bstream.stream = ctx->stream;
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_ordinal) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    e = cudaMalloc((void**) &ctx->d_flag, 4 * sizeof(int));      // [0] error flag, [1] task counter of persistent kernels
    if (e == cudaSuccess) e = cudaMemset(ctx->d_flag, 0, 4 * sizeof(int));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(flag) failed: %s", cudaGetErrorString(e));
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return GSB200_ERR_CUDA;
    }
    *out = ctx;
    return GSB200_OK;
}

void gsb200_destroy(gsb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& m : ctx->maps) m.release();
    for (auto& e : ctx->effects) e.release();
    for (auto& s : ctx->s_rows) s.release();
    ctx->s_io.release();
    for (auto& s : ctx->s_plan) s.release();
    for (auto& s : ctx->s_eval) s.release();
    for (auto& s : ctx->s_chain) s.release();
    for (auto& s : ctx->s_bv) s.release();
    ctx->s_null.release();
    ctx->s_select.release();
    ctx->local_sched.d_steps.release();
    ctx->local_sched.d_plan.release();
    ctx->local_sched.d_bstream.release();
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);        // the stream-ordered frees above
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->d_store) cudaFree(ctx->d_store);
    if (ctx->d_dict_sym) cudaFree(ctx->d_dict_sym);
    if (ctx->d_dict_n) cudaFree(ctx->d_dict_n);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int gsb200_synchronize(gsb200_ctx* ctx) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GSB200_OK;
}

void* gsb200_stream(gsb200_ctx* ctx) { return ctx ? (void*) ctx->stream : nullptr; }
uint64_t gsb200_kernel_launches(const gsb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

/* ------------------------------------------------------------------ maps */

int gsb200_map_set(gsb200_ctx* ctx, uint32_t map_index, uint32_t n_chr, const double* lambda,
                   const uint32_t* chr_offset, const double* dists, const uint32_t* marker_index,
                   const uint8_t* has_dists) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_REQUIRE(n_chr > 0 && lambda && chr_offset && marker_index && has_dists, GSB200_ERR_ARG,
                "gsb200_map_set: empty map or null array");
    GSB_REQUIRE(map_index < 4096, GSB200_ERR_ARG, "gsb200_map_set: map index %u too large", map_index);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t total = chr_offset[n_chr];
    GSB_REQUIRE(chr_offset[0] == 0, GSB200_ERR_ARG, "gsb200_map_set: chr_offset[0] must be 0");
    for (uint32_t c = 0; c < n_chr; ++c)
        GSB_REQUIRE(chr_offset[c] <= chr_offset[c + 1], GSB200_ERR_ARG, "gsb200_map_set: chr_offset not ascending");
    for (uint32_t p = 0; p < total; ++p)
        GSB_REQUIRE(marker_index[p] < ctx->n_markers, GSB200_ERR_ARG,
                    "gsb200_map_set: marker index %u out of range", marker_index[p]);

    std::vector<ChrInfo> v_chr(n_chr);
    std::vector<double> v_dists(std::max<uint32_t>(total, 1), 0.0);
    std::vector<uint32_t> v_mix(marker_index, marker_index + total);
    std::vector<uint32_t> v_bucket, v_chunk_chr, v_chunk_start;
    std::vector<uint8_t> covered(ctx->n_markers, 0);

    bool contiguous = true, disjoint = true;
    double lambda_sum = 0.0, lambda_max = 0.0;
    const double ninf = -std::numeric_limits<double>::infinity();
    for (uint32_t c = 0; c < n_chr; ++c) {
        const uint32_t lo = chr_offset[c], hi = chr_offset[c + 1];
        ChrInfo& ci = v_chr[c];
        ci.lambda = lambda[c];
        ci.exp_neg_lambda = lambda[c] > 0 ? std::exp(-lambda[c]) : 1.0;
        ci.len = hi - lo;
        ci.first = hi > lo ? marker_index[lo] : 0;
        ci.pos_off = lo;
        for (uint32_t p = lo; p < hi; ++p) {
            contiguous = contiguous && (marker_index[p] == ci.first + (p - lo));
            if (covered[marker_index[p]]) disjoint = false;
            covered[marker_index[p]] = 1;
        }
        if (lambda[c] > 0) { lambda_sum += lambda[c]; lambda_max = std::max(lambda_max, lambda[c]); }
        // The reference walks markers i = 0.. and advances through the sorted crossovers while
        // dists[i] > crossover (core.c:7902-7903); the number passed by marker i is therefore
        // the count of crossovers below the RUNNING MAXIMUM of dists[0..i] (NaN compares false,
        // so NaN never advances).  Storing that running maximum makes the strand at marker i
        // a pure function of #{crossovers < d[i]}, which is what the kernels evaluate.
        double run = ninf;
        bool any = false;
        for (uint32_t p = lo; p < hi; ++p) {
            double d = (has_dists[c] && dists) ? dists[p] : std::numeric_limits<double>::quiet_NaN();
            if (d > run) run = d;
            v_dists[p] = run;
            any = any || (run > ninf);
        }
        ci.searchable = any ? 1 : 0;
        // bucket index over [0,1): bucket b starts the search at #{i : !(d[i] > b / 2^shift)};
        // one extra entry (= len) closes the last bucket.
        uint32_t shift = 0;
        while ((1u << shift) < ci.len && shift < 20) ++shift;
        ci.bucket_shift = shift;
        ci.bucket_off = (uint32_t) v_bucket.size();
        const uint32_t nb = 1u << shift;
        uint32_t i = 0;
        for (uint32_t b = 0; b < nb; ++b) {
            const double edge = (double) b / (double) nb;
            while (i < ci.len && !(v_dists[lo + i] > edge)) ++i;
            v_bucket.push_back(i);
        }
        v_bucket.push_back(ci.len);
        for (uint32_t s0 = 0; s0 < ci.len; s0 += 32) { v_chunk_chr.push_back(c); v_chunk_start.push_back(s0); }
    }
    bool covers_all = true;
    for (uint32_t m = 0; m < ctx->n_markers; ++m) covers_all = covers_all && covered[m];
    const bool flat = contiguous && disjoint && covers_all;

    if (ctx->maps.size() <= map_index) ctx->maps.resize(map_index + 1);
    DeviceMap& m = ctx->maps[map_index];
    cudaStreamSynchronize(ctx->stream);
    m.release();
    m.n_chr = n_chr;
    m.n_positions = total;
    m.n_chunks = (uint32_t) v_chunk_chr.size();
    m.contiguous = contiguous;
    m.flat = flat;
    m.covers_all = covers_all;
    m.lambda_sum = lambda_sum;
    m.lambda_max = lambda_max;
    GSB_TRY(to_device(v_chr, &m.d_chr));
    GSB_TRY(to_device(v_dists, &m.d_dists));
    GSB_TRY(to_device(v_mix, &m.d_marker_index));
    GSB_TRY(to_device(v_bucket, &m.d_bucket));
    GSB_TRY(to_device(v_chunk_chr, &m.d_chunk_chr));
    GSB_TRY(to_device(v_chunk_start, &m.d_chunk_start));
    for (uint32_t mk = 0; mk < ctx->n_markers; ++mk) if (!covered[mk]) m.h_uncovered.push_back(mk);
    GSB_TRY(to_device(m.h_uncovered, &m.d_uncovered));
    m.present = true;
    return GSB200_OK;
}

int gsb200_map_delete(gsb200_ctx* ctx, uint32_t map_index) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_REQUIRE(map_index < ctx->maps.size() && ctx->maps[map_index].present, GSB200_ERR_ARG,
                "gsb200_map_delete: no map at index %u", map_index);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->maps[map_index].release();
    return GSB200_OK;
}

uint32_t gsb200_map_n_chr(const gsb200_ctx* ctx, uint32_t map_index) {
    if (!ctx || map_index >= ctx->maps.size() || !ctx->maps[map_index].present) return 0;
    return ctx->maps[map_index].n_chr;
}

/* ------------------------------------------------------------------ effects */

int gsb200_effects_set(gsb200_ctx* ctx, uint32_t eff_index, const uint32_t* cumn_alleles,
                       const char* allele, const double* eff, const double* centre) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_REQUIRE(cumn_alleles && allele && eff, GSB200_ERR_ARG, "gsb200_effects_set: null array");
    GSB_REQUIRE(eff_index < 65536, GSB200_ERR_ARG, "gsb200_effects_set: effect index %u too large", eff_index);
    GSB_CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t M = ctx->n_markers;
    for (uint32_t m = 1; m < M; ++m)
        GSB_REQUIRE(cumn_alleles[m - 1] <= cumn_alleles[m], GSB200_ERR_ARG, "gsb200_effects_set: cumn_alleles not ascending");
    const uint32_t nnz = cumn_alleles[M - 1];
    if (ctx->effects.size() <= eff_index) ctx->effects.resize(eff_index + 1);
    DeviceEffects& e = ctx->effects[eff_index];
    cudaStreamSynchronize(ctx->stream);
    e.release();
    e.cumn.assign(cumn_alleles, cumn_alleles + M);
    e.allele.assign((const uint8_t*) allele, (const uint8_t*) allele + nnz);
    e.eff.assign(eff, eff + nnz);
    e.has_centre = centre != nullptr;
    e.centre_sum = 0.0;
    if (centre) {
        e.centre.assign(centre, centre + M);
        for (uint32_t m = 0; m < M; ++m) e.centre_sum += centre[m];   // core.c:9604-9607 order
    }
    e.present = true;
    return GSB200_OK;
}

int gsb200_effects_delete(gsb200_ctx* ctx, uint32_t eff_index) {
    GSB_REQUIRE(ctx != nullptr, GSB200_ERR_ARG, "null context");
    GSB_REQUIRE(eff_index < ctx->effects.size() && ctx->effects[eff_index].present, GSB200_ERR_ARG,
                "gsb200_effects_delete: no effect set at index %u", eff_index);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->effects[eff_index].release();
    return GSB200_OK;
}

}  // extern "C"
