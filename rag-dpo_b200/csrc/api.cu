// api.cu — the C ABI of libb200rag.so (include/b200rag.h): handles, staging,
// stream/event plumbing and the orchestration of the kernels.  No CPU compute
// path exists here: without an sm_100 device every entry point fails.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include "kernels.h"

using namespace b200rag;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(e__ == cudaErrorMemoryAllocation ? RAG_ENOMEM : RAG_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                  \
    } while (0)

#define RAG_TRY(expr)             \
    do {                          \
        int r__ = (expr);         \
        if (r__ != RAG_OK) return r__; \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return RAG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = need + need / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(RAG_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        bytes = want;
        return RAG_OK;
    }
    template <typename T>
    T* as() { return reinterpret_cast<T*>(p); }
};

struct Ctx {
    bool inited = false;
    int device = -1;
    int sm_count = 0, cc_major = 0, cc_minor = 0, smem_optin = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[6] = {};
    bool ev_valid[6] = {};
    float timings[8] = {};
    int64_t n_launch = 0, n_fallback = 0;
    std::mutex mu;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    int32_t* pinned_small = nullptr;    // 4 KB for flags / counters read back inside a call
    // scratch (device)
    DevBuf q, allow, cand, cand_cnt, sample_keys, tau_keys, overflow, q16, q_resid, top, flags, tau, nflag, o_rows, o_scores, o_counts;
    DevBuf fb_q, fb_tau, fb_counts, fb_rows, fb_scores, fb_index;
    DevBuf bm_terms, bm_ranges, bm_cand, bm_rows, bm_scores, bm_counts, bm_allow, bm_scratch, bm_index;
    DevBuf rrf_ids, rrf_w, rrf_oi, rrf_os, rrf_oc;
    DevBuf stage_f32;
};
Ctx g;
int g_tc_min_batch = 2;     // smallest batch served by the tcgen05 path (env B200RAG_TC_MIN_BATCH)
int g_tc_b1_shadow = 1;     // batch-1 on fp32/fp16 corpora goes through the bf16 shadow (env B200RAG_TC_B1_SHADOW)
int g_exchange_timeout_ms = 10000;   // bound of the exchange's flag wait (option "exchange_timeout_ms")

int ensure_pinned(size_t need) {
    if (need <= g.pinned_bytes) return RAG_OK;
    if (g.pinned) cudaFreeHost(g.pinned);
    g.pinned = nullptr;
    g.pinned_bytes = 0;
    size_t want = need + need / 4 + 4096;
    cudaError_t e = cudaMallocHost(&g.pinned, want);
    if (e != cudaSuccess) return fail(RAG_ENOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    g.pinned_bytes = want;
    return RAG_OK;
}

int require_init() {
    if (!g.inited) return fail(RAG_ENODEV, "rag_init() has not succeeded: no sm_100 device bound (no CPU fallback)");
    cudaError_t e = cudaSetDevice(g.device);
    if (e != cudaSuccess) return fail(RAG_ENODEV, "cudaSetDevice(%d): %s", g.device, cudaGetErrorString(e));
    return RAG_OK;
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

void rec(int i) {
    if (cudaEventRecord(g.ev[i], g.stream) == cudaSuccess) g.ev_valid[i] = true;
}

}  // namespace

struct rag_corpus {
    int dim = 0, dtype = 0;
    int64_t cap = 0, n = 0;
    size_t row_bytes = 0;
    void* rows = nullptr;
    float* max_norm = nullptr;   // device scalar
    // bf16 shadow for the tensor-core path (fp32 / fp16 corpora), built lazily, dropped on mutation
    void* shadow = nullptr;
    int64_t shadow_cap = 0, shadow_rows = -1;
    float* shadow_resid = nullptr;
};

struct rag_bm25 {
    Bm25Device d{};
    std::vector<int64_t> h_term_ptr;
    std::vector<double> h_idf;
};

extern "C" {

const char* rag_last_error(void) { return g_err; }
int rag_abi_version(void) { return 1; }

int rag_init(int device) {
    std::lock_guard<std::mutex> lk(g.mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(RAG_ENODEV, "no CUDA device: %s (b200rag has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(RAG_EINVAL, "device %d out of range (0..%d)", device, count - 1);
    if (g.inited && g.device == device) return RAG_OK;
    if (g.inited) return fail(RAG_EINVAL, "already bound to device %d (one process per GPU)", g.device);
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RAG_ENODEV, "device %d is sm_%d%d; b200rag kernels are sm_100a only", device, prop.major, prop.minor);
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    g.cc_major = prop.major;
    g.cc_minor = prop.minor;
    g.smem_optin = (int)prop.sharedMemPerBlockOptin;
    CU_TRY(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    for (auto& ev : g.ev) CU_TRY(cudaEventCreate(&ev));
    CU_TRY(cudaMallocHost((void**)&g.pinned_small, 4096));
    if (const char* e = getenv("B200RAG_TC_MIN_BATCH")) g_tc_min_batch = std::max(1, atoi(e));
    if (const char* e = getenv("B200RAG_TC_B1_SHADOW")) g_tc_b1_shadow = atoi(e) != 0;
    g.inited = true;
    return RAG_OK;
}

int rag_set_stream(void* cuda_stream) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    g.stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : g.own_stream;
    return RAG_OK;
}

int rag_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (sm_count) *sm_count = g.sm_count;
    if (cc_major) *cc_major = g.cc_major;
    if (cc_minor) *cc_minor = g.cc_minor;
    size_t f = 0, t = 0;
    CU_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return RAG_OK;
}

int rag_set_option(const char* key, int64_t value) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!key) return fail(RAG_EINVAL, "key is NULL");
    if (!strcmp(key, "tc_min_batch")) g_tc_min_batch = (int)std::max<int64_t>(1, value);
    else if (!strcmp(key, "tc_b1_shadow")) g_tc_b1_shadow = value != 0;
    else if (!strcmp(key, "sample_div")) gemm_set_sample_div((int)value);
    else if (!strcmp(key, "balance_tail")) gemm_set_balance_tail((int)value);
    else if (!strcmp(key, "pair_mode")) gemm_set_pair_mode((int)value);
    else if (!strcmp(key, "sample_resident")) gemm_set_sample_resident((int)value);
    else if (!strcmp(key, "exchange_timeout_ms")) g_exchange_timeout_ms = (int)std::max<int64_t>(1, std::min<int64_t>(value, 600000));
    else return fail(RAG_EINVAL, "unknown option %s", key);
    return RAG_OK;
}

int rag_host_alloc(void** out, size_t bytes) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!out || bytes == 0) return fail(RAG_EINVAL, "bad host allocation request");
    CU_TRY(cudaMallocHost(out, bytes));
    return RAG_OK;
}

int rag_host_free(void* p) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (p && g.inited) cudaFreeHost(p);
    return RAG_OK;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int rag_last_timings(float* ms, int n) {
    std::lock_guard<std::mutex> lk(g.mu);
    for (int i = 0; i < n; ++i) ms[i] = i < 8 ? g.timings[i] : 0.f;
    return RAG_OK;
}

int rag_counters(int64_t* out, int n) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (n > 0) out[0] = g.n_launch;
    if (n > 1) out[1] = g.n_fallback;
    for (int i = 2; i < n; ++i) out[i] = 0;
    return RAG_OK;
}

// ---------------------------------------------------------------------------
// corpus
// ---------------------------------------------------------------------------
int rag_corpus_create(rag_corpus_t** out, int64_t capacity_rows, int dim, int dtype) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!out) return fail(RAG_EINVAL, "out is NULL");
    if (dtype != RAG_F32 && dtype != RAG_BF16 && dtype != RAG_F16) return fail(RAG_EINVAL, "bad dtype %d", dtype);
    if (dim <= 0 || dim % 64 != 0) return fail(RAG_ERANGE, "dim %d must be a positive multiple of 64", dim);
    if (dim > (dtype == RAG_F32 ? 1024 : 2048)) return fail(RAG_ERANGE, "dim %d too large for dtype %d", dim, dtype);
    if (capacity_rows < 0) return fail(RAG_EINVAL, "negative capacity");
    if (capacity_rows > 0x7FFFFFF0LL) return fail(RAG_ERANGE, "more than 2^31 rows per shard");
    rag_corpus* c = new rag_corpus();
    c->dim = dim;
    c->dtype = dtype;
    c->row_bytes = (size_t)dim * (dtype == RAG_F32 ? 4 : 2);
    c->cap = std::max<int64_t>(capacity_rows, 1);
    cudaError_t e = cudaMalloc(&c->rows, (size_t)c->cap * c->row_bytes);
    if (e != cudaSuccess) {
        const size_t want = (size_t)c->cap * c->row_bytes;
        delete c;
        return fail(RAG_ENOMEM, "cudaMalloc(%zu) for corpus failed: %s", want, cudaGetErrorString(e));
    }
    e = cudaMalloc(&c->max_norm, sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->max_norm, 0, sizeof(float));
    if (e != cudaSuccess) {
        cudaFree(c->rows);
        delete c;
        return fail(RAG_ECUDA, "corpus init: %s", cudaGetErrorString(e));
    }
    *out = c;
    return RAG_OK;
}

int rag_corpus_destroy(rag_corpus_t* c) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!c) return RAG_OK;
    if (g.inited) {
        cudaSetDevice(g.device);
        cudaStreamSynchronize(g.stream);
        cudaFree(c->rows);
        cudaFree(c->max_norm);
        cudaFree(c->shadow);
        cudaFree(c->shadow_resid);
    }
    delete c;
    return RAG_OK;
}

static int corpus_reserve_locked(rag_corpus* c, int64_t cap) {
    if (cap <= c->cap) return RAG_OK;
    if (cap > 0x7FFFFFF0LL) return fail(RAG_ERANGE, "more than 2^31 rows per shard");
    int64_t want = std::max(cap, c->cap + c->cap / 2);
    void* nr = nullptr;
    cudaError_t e = cudaMalloc(&nr, (size_t)want * c->row_bytes);
    if (e != cudaSuccess) {
        want = cap;
        e = cudaMalloc(&nr, (size_t)want * c->row_bytes);
    }
    if (e != cudaSuccess) return fail(RAG_ENOMEM, "cudaMalloc(%zu) growing corpus: %s", (size_t)want * c->row_bytes,
                                      cudaGetErrorString(e));
    CU_TRY(cudaMemcpyAsync(nr, c->rows, (size_t)c->n * c->row_bytes, cudaMemcpyDeviceToDevice, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    cudaFree(c->rows);
    c->rows = nr;
    c->cap = want;
    return RAG_OK;
}

int rag_corpus_reserve(rag_corpus_t* c, int64_t capacity_rows) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!c) return fail(RAG_EINVAL, "corpus is NULL");
    return corpus_reserve_locked(c, capacity_rows);
}

int rag_corpus_upload(rag_corpus_t* c, int64_t row0, int64_t nrows, const float* host_rows) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!c || (!host_rows && nrows > 0)) return fail(RAG_EINVAL, "NULL argument");
    if (row0 < 0 || nrows < 0 || row0 > c->n) return fail(RAG_EINVAL, "rows [%lld,+%lld) not contiguous with count %lld",
                                                          (long long)row0, (long long)nrows, (long long)c->n);
    if (nrows == 0) return RAG_OK;
    RAG_TRY(corpus_reserve_locked(c, row0 + nrows));
    const size_t chunk_rows = std::max<size_t>(1, (size_t)(32u << 20) / ((size_t)c->dim * 4));
    RAG_TRY(ensure_pinned(chunk_rows * c->dim * 4));
    if (c->dtype != RAG_F32) RAG_TRY(g.stage_f32.ensure(chunk_rows * c->dim * 4));
    for (int64_t r = 0; r < nrows; r += (int64_t)chunk_rows) {
        const int64_t nr = std::min<int64_t>((int64_t)chunk_rows, nrows - r);
        const size_t bytes = (size_t)nr * c->dim * 4;
        memcpy(g.pinned, host_rows + (size_t)r * c->dim, bytes);
        uint8_t* dst = reinterpret_cast<uint8_t*>(c->rows) + (size_t)(row0 + r) * c->row_bytes;
        if (c->dtype == RAG_F32) {
            CU_TRY(cudaMemcpyAsync(dst, g.pinned, bytes, cudaMemcpyHostToDevice, g.stream));
        } else {
            CU_TRY(cudaMemcpyAsync(g.stage_f32.p, g.pinned, bytes, cudaMemcpyHostToDevice, g.stream));
            CU_TRY(convert_rows_launch(g.stage_f32.as<float>(), dst, c->dtype, nr * c->dim, g.stream));
            ++g.n_launch;
        }
        CU_TRY(row_norm_max_launch(dst, c->dtype, nr, c->dim, c->max_norm, g.stream));
        ++g.n_launch;
        CU_TRY(cudaStreamSynchronize(g.stream));      // the pinned chunk is reused
    }
    c->n = std::max(c->n, row0 + nrows);
    c->shadow_rows = -1;
    return RAG_OK;
}

int rag_corpus_download(const rag_corpus_t* c, int64_t row0, int64_t nrows, float* host_rows) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!c || (!host_rows && nrows > 0)) return fail(RAG_EINVAL, "NULL argument");
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->n) return fail(RAG_EINVAL, "rows out of range");
    if (nrows == 0) return RAG_OK;
    const size_t chunk_rows = std::max<size_t>(1, (size_t)(32u << 20) / ((size_t)c->dim * 4));
    RAG_TRY(ensure_pinned(chunk_rows * c->dim * 4));
    RAG_TRY(g.stage_f32.ensure(chunk_rows * c->dim * 4));
    for (int64_t r = 0; r < nrows; r += (int64_t)chunk_rows) {
        const int64_t nr = std::min<int64_t>((int64_t)chunk_rows, nrows - r);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(c->rows) + (size_t)(row0 + r) * c->row_bytes;
        CU_TRY(widen_rows_launch(src, c->dtype, g.stage_f32.as<float>(), nr * c->dim, g.stream));
        ++g.n_launch;
        CU_TRY(cudaMemcpyAsync(g.pinned, g.stage_f32.p, (size_t)nr * c->dim * 4, cudaMemcpyDeviceToHost, g.stream));
        CU_TRY(cudaStreamSynchronize(g.stream));
        memcpy(host_rows + (size_t)r * c->dim, g.pinned, (size_t)nr * c->dim * 4);
    }
    return RAG_OK;
}

int rag_corpus_compact(rag_corpus_t* c, const int64_t* keep_rows, int64_t nkeep) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!c || (!keep_rows && nkeep > 0) || nkeep < 0 || nkeep > c->n) return fail(RAG_EINVAL, "bad compact arguments");
    for (int64_t i = 0; i < nkeep; ++i) {
        if (keep_rows[i] < 0 || keep_rows[i] >= c->n || (i > 0 && keep_rows[i] <= keep_rows[i - 1]))
            return fail(RAG_EINVAL, "keep_rows must be strictly ascending rows below the count");
    }
    if (nkeep == c->n) return RAG_OK;
    void* nr = nullptr;
    const int64_t ncap = std::max<int64_t>(nkeep, 1);
    CU_TRY(cudaMalloc(&nr, (size_t)ncap * c->row_bytes));
    if (nkeep > 0) {
        DevBuf keep;
        int rc = keep.ensure((size_t)nkeep * 8);
        if (rc != RAG_OK) { cudaFree(nr); return rc; }
        CU_TRY(cudaMemcpyAsync(keep.p, keep_rows, (size_t)nkeep * 8, cudaMemcpyHostToDevice, g.stream));
        CU_TRY(gather_rows_launch(c->rows, nr, keep.as<int64_t>(), nkeep, (int)c->row_bytes, g.stream));
        ++g.n_launch;
        CU_TRY(cudaStreamSynchronize(g.stream));
        cudaFree(keep.p);
    }
    cudaFree(c->rows);
    c->rows = nr;
    c->cap = ncap;
    c->n = nkeep;
    c->shadow_rows = -1;
    // the max norm only ever over-estimates after a delete, which keeps the bound valid
    return RAG_OK;
}

int rag_corpus_count(const rag_corpus_t* c, int64_t* n) {
    if (!c || !n) return fail(RAG_EINVAL, "NULL argument");
    *n = c->n;
    return RAG_OK;
}

int rag_corpus_fill_synthetic(rag_corpus_t* c, uint64_t seed, int64_t gen_row0, int64_t row0, int64_t nrows) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!c) return fail(RAG_EINVAL, "corpus is NULL");
    if (row0 < 0 || nrows < 0 || row0 > c->n) return fail(RAG_EINVAL, "rows not contiguous with count");
    if (nrows == 0) return RAG_OK;
    RAG_TRY(corpus_reserve_locked(c, row0 + nrows));
    CU_TRY(fill_synthetic_launch(c->rows, c->dtype, row0, nrows, c->dim, seed, gen_row0, g.stream));
    uint8_t* dst = reinterpret_cast<uint8_t*>(c->rows) + (size_t)row0 * c->row_bytes;
    CU_TRY(row_norm_max_launch(dst, c->dtype, nrows, c->dim, c->max_norm, g.stream));
    g.n_launch += 2;
    CU_TRY(cudaStreamSynchronize(g.stream));
    c->n = std::max(c->n, row0 + nrows);
    c->shadow_rows = -1;
    return RAG_OK;
}

int rag_corpus_device_ptr(const rag_corpus_t* c, void** rows_dev) {
    if (!c || !rows_dev) return fail(RAG_EINVAL, "NULL argument");
    *rows_dev = c->rows;
    return RAG_OK;
}

// ---------------------------------------------------------------------------
// dense top-k
// ---------------------------------------------------------------------------
// filter error bound of the fp32 CUDA-core scan, relative to |q|*max|x|:
// <= 16 chained FMAs + 2 + 5 tree levels, each 2^-24 -> < 2^-19.
static const double kEpsScan = 1.0 / 524288.0;

// accumulation error of the tensor-core filter relative to |q|*max|x|: 64 chained K=16 blocks plus the
// in-block tree, fp32 accumulators (truncating) -> < 70 * 2^-23 < 2^-16.
static const double kEpsTc = 1.0 / 65536.0;

// 16-bit operand of the tensor-core path: bf16 / fp16 rows as stored, or a lazily built bf16 shadow of fp32 rows
static int ensure_bf16_operand(rag_corpus* c, const void** x16, const float** x_resid) {
    if (c->dtype == RAG_BF16 || c->dtype == RAG_F16) {
        *x16 = c->rows;
        *x_resid = nullptr;
        return RAG_OK;
    }
    if (c->shadow_rows != c->n) {
        if (c->shadow_cap < c->n) {
            if (c->shadow) cudaFree(c->shadow);
            c->shadow = nullptr;
            c->shadow_cap = 0;
            cudaError_t e = cudaMalloc(&c->shadow, (size_t)c->cap * c->dim * 2);
            if (e != cudaSuccess)
                return fail(RAG_ENOMEM, "cudaMalloc(%zu) for the bf16 shadow: %s", (size_t)c->cap * c->dim * 2,
                            cudaGetErrorString(e));
            c->shadow_cap = c->cap;
        }
        if (!c->shadow_resid) {
            CU_TRY(cudaMalloc((void**)&c->shadow_resid, sizeof(float)));
        }
        CU_TRY(cudaMemsetAsync(c->shadow_resid, 0, sizeof(float), g.stream));
        CU_TRY(shadow_launch(c->rows, c->dtype, c->n, c->dim, c->shadow, c->shadow_resid, g.stream));
        ++g.n_launch;
        c->shadow_rows = c->n;
    }
    *x16 = c->shadow;
    *x_resid = c->shadow_resid;
    return RAG_OK;
}

static int dense_core_slice(rag_corpus* c, const float* q_dev, int B, int k, const uint8_t* allow_dev,
                            int32_t* o_rows, double* o_scores, int32_t* o_counts, bool defer);
static int dense_finish(rag_corpus* c, const float* q_dev, int B, int k, const uint8_t* allow_dev, int32_t* o_rows,
                        double* o_scores, int32_t* o_counts, bool* redone);

// batches larger than one contraction launch serves are processed in slices
static int dense_core(rag_corpus* c, const float* q_dev, int B, int k, const uint8_t* allow_dev, int32_t* o_rows,
                      double* o_scores, int32_t* o_counts) {
    const int step = gemm_max_batch();
    if (B <= step) return dense_core_slice(c, q_dev, B, k, allow_dev, o_rows, o_scores, o_counts, false);
    float acc[8] = {};
    for (int off = 0; off < B; off += step) {
        const int nb = std::min(step, B - off);
        RAG_TRY(dense_core_slice(c, q_dev + (size_t)off * c->dim, nb, k, allow_dev, o_rows + (size_t)off * k,
                                 o_scores + (size_t)off * k, o_counts + off, false));
        for (int i = 0; i < 8; ++i) acc[i] += g.timings[i];
    }
    for (int i = 0; i < 8; ++i) g.timings[i] = acc[i];
    return RAG_OK;
}

static int dense_core_slice(rag_corpus* c, const float* q_dev, int B, int k, const uint8_t* allow_dev,
                            int32_t* o_rows, double* o_scores, int32_t* o_counts, bool defer) {
    // tensor-core path: every batch >= g_tc_min_batch (default 2); a single query only on an fp32 corpus (the
    // filter then streams the bf16 shadow: half the bytes of the fp32 rows) that is large enough to matter
    const bool use_tc = c->n > 0 && (B >= g_tc_min_batch || (g_tc_b1_shadow && c->dtype == RAG_F32 && c->n >= 262144));
    const int kp = use_tc ? std::max(64, next_pow2(2 * k + 1)) : std::max(16, next_pow2(k + 6));
    for (auto& v : g.ev_valid) v = false;
    for (auto& t : g.timings) t = 0.f;

    RAG_TRY(g.top.ensure((size_t)B * kp * 8));
    RAG_TRY(g.flags.ensure((size_t)B * 4));
    RAG_TRY(g.tau.ensure((size_t)B * 4));
    RAG_TRY(g.nflag.ensure(4));
    CU_TRY(cudaMemsetAsync(g.nflag.p, 0, 4, g.stream));
    const float* q_resid = nullptr;
    const float* x_resid = nullptr;
    const uint64_t* tau_keys = nullptr;
    int32_t* overflow = nullptr;
    // inputs of the fused merge + refine launch
    const int32_t* m_counts = nullptr;
    int m_flat = 0, m_lists = 0, m_len = 0;

    if (c->n == 0) {
        CU_TRY(cudaMemsetAsync(g.top.p, 0, (size_t)B * kp * 8, g.stream));
    } else if (use_tc) {
        // ---- tcgen05 contraction + fused top-k (dense_gemm.cu)
        const void* x16 = nullptr;
        RAG_TRY(ensure_bf16_operand(c, &x16, &x_resid));
        GemmParams p{};
        p.n_rows = c->n;
        p.dim = c->dim;
        p.n_queries = B;
        p.kp = kp;
        p.allow = allow_dev;
        p.fp16_operands = c->dtype == RAG_F16;
        int grid = 0;
        const size_t smem = gemm_plan(p, g.sm_count, g.smem_optin, &grid);
        if (smem == 0) return fail(RAG_ERANGE, "k=%d does not fit the contraction kernel's shared memory", k);
        const int bpad = gemm_padded_queries(B);
        RAG_TRY(g.q16.ensure((size_t)bpad * c->dim * 2));
        RAG_TRY(g.q_resid.ensure((size_t)bpad * 4));
        const int sm = gemm_sample_m();
        RAG_TRY(g.cand.ensure((size_t)bpad * p.list_cap * 8));
        RAG_TRY(g.cand_cnt.ensure((size_t)bpad * 4));
        RAG_TRY(g.sample_keys.ensure((size_t)bpad * p.n_lists * sm * 8));
        RAG_TRY(g.tau_keys.ensure((size_t)bpad * sm * 8));
        RAG_TRY(g.overflow.ensure((size_t)bpad * 4));
        p.cand = g.cand.as<uint64_t>();
        p.cand_cnt = g.cand_cnt.as<int32_t>();
        p.sample_keys = g.sample_keys.as<uint64_t>();
        p.tau_keys = p.use_sample ? g.tau_keys.as<uint64_t>() : nullptr;
        CU_TRY(query_prep_launch(q_dev, B, bpad, c->dim, g.q16.p, g.q_resid.as<float>(), p.fp16_operands, g.stream));
        ++g.n_launch;
        q_resid = g.q_resid.as<float>();
        CU_TRY(cudaMemsetAsync(g.cand_cnt.p, 0, (size_t)bpad * 4, g.stream));
        rec(0);
        if (p.use_sample) {
            // sample pass -> per-query threshold (the 8th best sample score)
            CU_TRY(gemm_launch(p, 0, g.q16.p, x16, grid, smem, g.stream));
            CU_TRY(sample_tau_launch(g.sample_keys.as<uint64_t>(), B, p.n_lists * sm, sm, g.tau_keys.as<uint64_t>(),
                                     g.stream));
            g.n_launch += 2;
        }
        rec(5);                         // timings[6]: the main pass alone (the launch the roofline is quoted on)
        CU_TRY(gemm_launch(p, 1, g.q16.p, x16, grid, smem, g.stream));
        ++g.n_launch;
        rec(1);
        // the per-query list is one contiguous block: the 8 merge warps split it (flat count per query)
        m_counts = g.cand_cnt.as<int32_t>();
        m_flat = 1;
        m_lists = 8;
        m_len = p.list_cap / 8;
        tau_keys = p.tau_keys;
        overflow = g.overflow.as<int32_t>();
    } else {
        // ---- CUDA-core scan (dense_scan.cu): one launch per <= 4 queries
        ScanParams p{};
        p.rows = c->rows;
        p.n_rows = c->n;
        p.dim = c->dim;
        p.kp = kp;
        p.mode = 0;
        p.allow = allow_dev;
        const int n_groups = (B + 3) / 4;
        // every group of the call uses the template width of the first one, so the
        // consumer-warp count (hence n_lists and the cand layout) is uniform
        p.n_queries = std::min(B, 4);
        p.nq_t = scan_nq_template(p.n_queries);
        int grid = 0, nch = 0;
        const size_t smem = scan_plan(p, c->dtype, g.sm_count, g.smem_optin, &grid, &nch);
        if (smem == 0) return fail(RAG_ERANGE, "k=%d does not fit the scan kernel's shared memory", k);
        const int n_lists = p.n_lists;
        RAG_TRY(g.cand.ensure((size_t)B * n_lists * kp * 8));
        rec(0);
        for (int gi = 0; gi < n_groups; ++gi) {
            ScanParams pg = p;
            pg.n_queries = std::min(4, B - gi * 4);
            pg.q = q_dev + (size_t)gi * 4 * c->dim;
            pg.cand = g.cand.as<uint64_t>() + (size_t)gi * 4 * n_lists * kp;
            CU_TRY(scan_launch(pg, c->dtype, nch, grid, smem, g.stream));
            ++g.n_launch;
        }
        rec(1);
        m_lists = n_lists;
        m_len = kp;
    }
    rec(2);
    RefineParams rp{};
    rp.top = g.top.as<uint64_t>();
    rp.rows = c->rows;
    rp.q = q_dev;
    rp.dtype = c->dtype;
    rp.dim = c->dim;
    rp.kp = kp;
    rp.k = k;
    rp.B = B;
    rp.eps_rel = use_tc ? kEpsTc : kEpsScan;
    rp.q_resid = q_resid;
    rp.x_resid = x_resid;
    rp.tau_keys = tau_keys;
    rp.tau_stride = gemm_sample_m();
    rp.max_row_norm = c->max_norm;
    rp.out_rows = o_rows;
    rp.out_scores = o_scores;
    rp.out_counts = o_counts;
    rp.flags = g.flags.as<int32_t>();
    rp.tau = g.tau.as<float>();
    rp.n_flagged = g.nflag.as<int32_t>();
    if (c->n == 0) {
        CU_TRY(refine_launch(rp, g.stream));
    } else {
        CU_TRY(merge_refine_launch(g.cand.as<uint64_t>(), m_counts, m_flat, m_lists, m_len, use_tc ? 0 : 1, overflow, rp,
                                   g.stream));
    }
    ++g.n_launch;
    rec(3);

    CU_TRY(cudaMemcpyAsync(g.pinned_small, g.nflag.p, 4, cudaMemcpyDeviceToHost, g.stream));
    if (defer) return RAG_OK;          // the caller syncs once (results + flag) and then calls dense_finish
    CU_TRY(cudaStreamSynchronize(g.stream));
    return dense_finish(c, q_dev, B, k, allow_dev, o_rows, o_scores, o_counts, nullptr);
}

// after the stream was synchronised: run the exact fallback pass for flagged queries (rare), fill the timings.
// *redone is set when the fallback rewrote some results.
static int dense_finish(rag_corpus* c, const float* q_dev, int B, int k, const uint8_t* allow_dev, int32_t* o_rows,
                        double* o_scores, int32_t* o_counts, bool* redone) {
    const int n_flagged = *g.pinned_small;
    if (redone) *redone = n_flagged > 0;

    if (n_flagged > 0) {
        // ---- fallback: the margin check failed for some queries (near-ties deeper than
        // the candidate list).  Collect every row whose filter score can still reach
        // the exact k-th score, re-score all of them exactly, select.
        ++g.n_fallback;
        std::vector<int32_t> h_flags(B);
        std::vector<float> h_tau(B);
        CU_TRY(cudaMemcpyAsync(h_flags.data(), g.flags.p, (size_t)B * 4, cudaMemcpyDeviceToHost, g.stream));
        CU_TRY(cudaMemcpyAsync(h_tau.data(), g.tau.p, (size_t)B * 4, cudaMemcpyDeviceToHost, g.stream));
        CU_TRY(cudaStreamSynchronize(g.stream));
        std::vector<int32_t> idx;
        for (int b = 0; b < B; ++b)
            if (h_flags[b]) idx.push_back(b);
        const int nf = (int)idx.size();
        RAG_TRY(g.fb_q.ensure((size_t)nf * c->dim * 4));
        RAG_TRY(g.fb_tau.ensure((size_t)nf * 4));
        RAG_TRY(g.fb_counts.ensure((size_t)nf * 4));
        RAG_TRY(g.fb_index.ensure((size_t)nf * 4));
        std::vector<float> tau_c(nf);
        for (int i = 0; i < nf; ++i) {
            tau_c[i] = h_tau[idx[i]];
            CU_TRY(cudaMemcpyAsync(g.fb_q.as<float>() + (size_t)i * c->dim, q_dev + (size_t)idx[i] * c->dim,
                                   (size_t)c->dim * 4, cudaMemcpyDeviceToDevice, g.stream));
        }
        CU_TRY(cudaMemcpyAsync(g.fb_tau.p, tau_c.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, g.stream));
        CU_TRY(cudaMemcpyAsync(g.fb_index.p, idx.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, g.stream));
        int cap = 4096;
        std::vector<unsigned> h_counts(nf);
        for (int attempt = 0; attempt < 2; ++attempt) {
            RAG_TRY(g.fb_rows.ensure((size_t)nf * cap * 4));
            RAG_TRY(g.fb_scores.ensure((size_t)nf * cap * 8));
            CU_TRY(cudaMemsetAsync(g.fb_counts.p, 0, (size_t)nf * 4, g.stream));
            for (int gi = 0; gi * 4 < nf; ++gi) {
                ScanParams p{};
                p.rows = c->rows;
                p.n_rows = c->n;
                p.dim = c->dim;
                p.kp = 16;
                p.mode = 1;
                p.allow = allow_dev;
                p.n_queries = std::min(4, nf - gi * 4);
                p.nq_t = scan_nq_template(p.n_queries);
                p.q = g.fb_q.as<float>() + (size_t)gi * 4 * c->dim;
                p.tau = g.fb_tau.as<float>() + gi * 4;
                p.collect_count = g.fb_counts.as<unsigned>() + gi * 4;
                p.collect_rows = g.fb_rows.as<uint32_t>() + (size_t)gi * 4 * cap;
                p.collect_cap = cap;
                int grid = 0, nch = 0;
                size_t smem = scan_plan(p, c->dtype, g.sm_count, g.smem_optin, &grid, &nch);
                if (smem == 0) return fail(RAG_ERANGE, "fallback scan does not fit shared memory");
                CU_TRY(scan_launch(p, c->dtype, nch, grid, smem, g.stream));
                ++g.n_launch;
            }
            CU_TRY(cudaMemcpyAsync(h_counts.data(), g.fb_counts.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, g.stream));
            CU_TRY(cudaStreamSynchronize(g.stream));
            unsigned mx = 0;
            for (unsigned v : h_counts) mx = std::max(mx, v);
            if (mx <= (unsigned)cap) break;
            if (attempt == 1) return fail(RAG_ECUDA, "fallback collect overflowed twice (%u > %d)", mx, cap);
            cap = (int)mx + 64;
        }
        CollectSelectParams sp{};
        sp.rows_list = g.fb_rows.as<uint32_t>();
        sp.counts = g.fb_counts.as<unsigned>();
        sp.cap = cap;
        sp.query_index = g.fb_index.as<int32_t>();
        sp.rows = c->rows;
        sp.q = g.fb_q.as<float>();
        sp.dtype = c->dtype;
        sp.dim = c->dim;
        sp.k = k;
        sp.nq = nf;
        sp.scratch_scores = g.fb_scores.as<double>();
        sp.out_rows = o_rows;
        sp.out_scores = o_scores;
        sp.out_counts = o_counts;
        CU_TRY(collect_select_launch(sp, g.stream));
        ++g.n_launch;
        rec(4);
        CU_TRY(cudaStreamSynchronize(g.stream));
    }
    // stage timings (events are complete: the stream was synchronised)
    auto el = [&](int a, int b) {
        float ms = 0.f;
        if (g.ev_valid[a] && g.ev_valid[b] && cudaEventElapsedTime(&ms, g.ev[a], g.ev[b]) == cudaSuccess) return ms;
        return 0.f;
    };
    g.timings[0] = el(0, 1);
    g.timings[1] = el(1, 2);
    g.timings[2] = el(2, 3);
    g.timings[3] = n_flagged > 0 ? el(3, 4) : 0.f;
    g.timings[6] = el(5, 1);
    return RAG_OK;
}

static int dense_check(rag_corpus* c, const void* q, int B, int k, const void* o_rows, const void* o_scores,
                       const void* o_counts) {
    if (!c || !q || !o_rows || !o_scores || !o_counts) return fail(RAG_EINVAL, "NULL argument");
    if (B <= 0) return fail(RAG_EINVAL, "B must be positive");
    if (k <= 0 || k > RAG_MAX_K) return fail(RAG_ERANGE, "k=%d outside 1..%d", k, RAG_MAX_K);
    return RAG_OK;
}

int rag_dense_topk_dev(rag_corpus_t* c, const float* q_dev, int B, int k, const uint8_t* allow_bitmap_dev,
                       int32_t* out_rows_dev, double* out_scores_dev, int32_t* out_counts_dev) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    RAG_TRY(dense_check(c, q_dev, B, k, out_rows_dev, out_scores_dev, out_counts_dev));
    return dense_core(c, q_dev, B, k, allow_bitmap_dev, out_rows_dev, out_scores_dev, out_counts_dev);
}

int rag_dense_topk(rag_corpus_t* c, const float* q, int B, int k, const uint8_t* allow_bitmap, int32_t* out_rows,
                   double* out_scores, int32_t* out_counts) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    RAG_TRY(dense_check(c, q, B, k, out_rows, out_scores, out_counts));
    const auto t_enter = std::chrono::steady_clock::now();
    float ms_queued = 0.f;
    const size_t qb = (size_t)B * c->dim * 4;
    const size_t ab = allow_bitmap ? (size_t)((c->n + 7) / 8) : 0;
    const size_t rb = (size_t)B * k * 4, sb = (size_t)B * k * 8, cb = (size_t)B * 4;
    RAG_TRY(g.q.ensure(qb));
    RAG_TRY(g.o_rows.ensure(rb));
    RAG_TRY(g.o_scores.ensure(sb));
    RAG_TRY(g.o_counts.ensure(cb));
    if (ab) RAG_TRY(g.allow.ensure(ab + 16));
    // inputs: DMA straight from page-locked caller memory, otherwise stage through the pinned block
    const bool q_pinned = is_pinned(q);
    const bool out_pinned = is_pinned(out_scores) && is_pinned(out_rows) && is_pinned(out_counts);
    RAG_TRY(ensure_pinned(std::max(qb + ab, sb + rb + cb)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(g.pinned);
    if (q_pinned) {
        CU_TRY(cudaMemcpyAsync(g.q.p, q, qb, cudaMemcpyHostToDevice, g.stream));
    } else {
        memcpy(pin, q, qb);
        CU_TRY(cudaMemcpyAsync(g.q.p, pin, qb, cudaMemcpyHostToDevice, g.stream));
    }
    if (ab) {
        memcpy(pin + qb, allow_bitmap, ab);
        CU_TRY(cudaMemcpyAsync(g.allow.p, pin + qb, ab, cudaMemcpyHostToDevice, g.stream));
    }
    const uint8_t* allow_dev = ab ? g.allow.as<uint8_t>() : nullptr;
    const bool one_sync = B <= gemm_max_batch();
    if (one_sync) {
        // launch everything, read results and the fallback flag back with ONE synchronisation
        RAG_TRY(dense_core_slice(c, g.q.as<float>(), B, k, allow_dev, g.o_rows.as<int32_t>(), g.o_scores.as<double>(),
                                 g.o_counts.as<int32_t>(), true));
    } else {
        RAG_TRY(dense_core(c, g.q.as<float>(), B, k, allow_dev, g.o_rows.as<int32_t>(), g.o_scores.as<double>(),
                           g.o_counts.as<int32_t>()));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        uint8_t* ds = out_pinned ? reinterpret_cast<uint8_t*>(out_scores) : pin;
        uint8_t* dr = out_pinned ? reinterpret_cast<uint8_t*>(out_rows) : pin + sb;
        uint8_t* dc = out_pinned ? reinterpret_cast<uint8_t*>(out_counts) : pin + sb + rb;
        CU_TRY(cudaMemcpyAsync(ds, g.o_scores.p, sb, cudaMemcpyDeviceToHost, g.stream));
        CU_TRY(cudaMemcpyAsync(dr, g.o_rows.p, rb, cudaMemcpyDeviceToHost, g.stream));
        CU_TRY(cudaMemcpyAsync(dc, g.o_counts.p, cb, cudaMemcpyDeviceToHost, g.stream));
        if (attempt == 0)
            ms_queued = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
        CU_TRY(cudaStreamSynchronize(g.stream));
        bool redone = false;
        if (one_sync && attempt == 0)
            RAG_TRY(dense_finish(c, g.q.as<float>(), B, k, allow_dev, g.o_rows.as<int32_t>(), g.o_scores.as<double>(),
                                 g.o_counts.as<int32_t>(), &redone));
        if (!redone) break;            // otherwise copy the rewritten results once more
    }
    if (!out_pinned) {
        memcpy(out_scores, pin, sb);
        memcpy(out_rows, pin + sb, rb);
        memcpy(out_counts, pin + sb + rb, cb);
    }
    // host-side view of the call: time spent queueing work (before the one synchronisation) and in total
    g.timings[4] = ms_queued;
    g.timings[5] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return RAG_OK;
}

// ---------------------------------------------------------------------------
// peer-memory exchange (exchange.cu)
// ---------------------------------------------------------------------------
struct rag_exchange {
    int world = 0, rank = 0;
    size_t slot_bytes = 0, total_bytes = 0;
    uint8_t* local = nullptr;                 // payload[2][world][slot] | flags[2][world] | counter
    void* mapped[kMaxExchangeRanks] = {};     // peers' buffers as opened in this process (mine: == local)
    bool connected = false;
    uint64_t epoch = 0;
    cudaStream_t stream = nullptr;            // every step runs on the stream of the first one
    ExchangeDev dev{};
};

static size_t exchange_flags_offset(const rag_exchange* ex) { return 2 * (size_t)ex->world * ex->slot_bytes; }

int rag_exchange_create(rag_exchange_t** out, int world, int rank, size_t slot_bytes, void* handle_out) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!out || !handle_out) return fail(RAG_EINVAL, "NULL argument");
    if (world < 1 || world > kMaxExchangeRanks || rank < 0 || rank >= world || slot_bytes == 0)
        return fail(RAG_ERANGE, "world=%d rank=%d slot_bytes=%zu outside the supported range", world, rank, slot_bytes);
    static_assert(sizeof(cudaIpcMemHandle_t) == RAG_IPC_HANDLE_BYTES, "handle size");
    rag_exchange* ex = new rag_exchange();
    ex->world = world;
    ex->rank = rank;
    ex->slot_bytes = (slot_bytes + 255) / 256 * 256;
    ex->total_bytes = exchange_flags_offset(ex) + 2 * (size_t)world * 8 + 256;
    cudaError_t e = cudaMalloc((void**)&ex->local, ex->total_bytes);
    if (e != cudaSuccess) {
        const size_t want = ex->total_bytes;
        delete ex;
        cudaGetLastError();
        return fail(RAG_ENOMEM, "cudaMalloc(%zu) for the exchange buffer: %s", want, cudaGetErrorString(e));
    }
    e = cudaMemset(ex->local, 0, ex->total_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();      // zeroed flags are in place before any peer can map them
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ex->local);
    if (e != cudaSuccess) {
        cudaFree(ex->local);
        delete ex;
        cudaGetLastError();
        return fail(RAG_ECUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, sizeof(h));
    *out = ex;
    return RAG_OK;
}

int rag_exchange_connect(rag_exchange_t* ex, const void* handles) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!ex || !handles) return fail(RAG_EINVAL, "NULL argument");
    if (ex->connected) return fail(RAG_EINVAL, "exchange already connected");
    const size_t fo = exchange_flags_offset(ex);
    for (int r = 0; r < ex->world; ++r) {
        void* p = ex->local;
        if (r != ex->rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, reinterpret_cast<const uint8_t*>(handles) + (size_t)r * sizeof(h), sizeof(h));
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int q = 0; q < r; ++q)
                    if (q != ex->rank && ex->mapped[q]) { cudaIpcCloseMemHandle(ex->mapped[q]); ex->mapped[q] = nullptr; }
                return fail(RAG_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            }
        }
        ex->mapped[r] = p;
        ex->dev.peer_base[r] = reinterpret_cast<uint8_t*>(p);
        ex->dev.peer_flags[r] = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(p) + fo);
    }
    ex->dev.my_base = ex->local;
    ex->dev.my_flags = reinterpret_cast<uint64_t*>(ex->local + fo);
    ex->dev.done_counter = reinterpret_cast<unsigned*>(ex->local + fo + 2 * (size_t)ex->world * 8);
    ex->dev.world = ex->world;
    ex->dev.rank = ex->rank;
    ex->dev.slot_bytes = ex->slot_bytes;
    ex->connected = true;
    return RAG_OK;
}

int rag_exchange_destroy(rag_exchange_t* ex) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!ex) return RAG_OK;
    if (g.inited) {
        cudaDeviceSynchronize();
        for (int r = 0; r < ex->world; ++r)
            if (r != ex->rank && ex->mapped[r]) cudaIpcCloseMemHandle(ex->mapped[r]);
        if (ex->local) cudaFree(ex->local);
        cudaGetLastError();
    }
    delete ex;
    return RAG_OK;
}

static int exchange_step(rag_exchange_t* ex, const double* my_scores_dev, const int64_t* my_ids_dev,
                         const int32_t* my_rows_dev, int64_t row_lo, int B, int k, double* out_scores_dev,
                         int64_t* out_ids_dev, int32_t* out_counts_dev) {
    RAG_TRY(require_init());
    if (!ex || !my_scores_dev || (!my_ids_dev && !my_rows_dev) || !out_scores_dev || !out_ids_dev || !out_counts_dev)
        return fail(RAG_EINVAL, "NULL argument");
    if (!ex->connected) return fail(RAG_EINVAL, "exchange not connected");
    if (B <= 0 || k <= 0 || k > RAG_MAX_K || (int64_t)ex->world * k > 8192)
        return fail(RAG_ERANGE, "world=%d B=%d k=%d outside the supported range", ex->world, B, k);
    if ((size_t)B * k * 16 > ex->slot_bytes)
        return fail(RAG_ERANGE, "B*k*16 = %zu bytes exceed the exchange slot (%zu)", (size_t)B * k * 16, ex->slot_bytes);
    // the two buffer parities are only safe when every step of this exchange is ordered on ONE stream
    if (ex->epoch == 0) ex->stream = g.stream;
    else if (ex->stream != g.stream)
        return fail(RAG_EINVAL, "the exchange is pinned to the stream of its first step (rag_set_stream changed it)");
    ex->dev.timeout_cycles = (long long)g_exchange_timeout_ms * 2000000LL;
    // the epoch advances only once the push is queued: a failed launch must not desynchronise the ranks
    CU_TRY(exchange_push_launch(ex->dev, my_scores_dev, my_ids_dev, my_rows_dev, row_lo, B, k, ex->epoch + 1, g.stream));
    ++ex->epoch;
    ++g.n_launch;
    CU_TRY(exchange_merge_launch(ex->dev, B, k, ex->epoch, out_scores_dev, out_ids_dev, out_counts_dev, g.stream));
    ++g.n_launch;
    return RAG_OK;
}

int rag_exchange_merge_topk_dev(rag_exchange_t* ex, const double* my_scores_dev, const int64_t* my_ids_dev, int B, int k,
                                double* out_scores_dev, int64_t* out_ids_dev, int32_t* out_counts_dev) {
    std::lock_guard<std::mutex> lk(g.mu);
    return exchange_step(ex, my_scores_dev, my_ids_dev, nullptr, 0, B, k, out_scores_dev, out_ids_dev, out_counts_dev);
}

int rag_exchange_merge_rows_dev(rag_exchange_t* ex, const double* my_scores_dev, const int32_t* my_rows_dev,
                                int64_t row_lo, int B, int k, double* out_scores_dev, int64_t* out_ids_dev,
                                int32_t* out_counts_dev) {
    std::lock_guard<std::mutex> lk(g.mu);
    return exchange_step(ex, my_scores_dev, nullptr, my_rows_dev, row_lo, B, k, out_scores_dev, out_ids_dev,
                         out_counts_dev);
}

int rag_exchange_status(rag_exchange_t* ex, int* timed_out) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!ex || !timed_out) return fail(RAG_EINVAL, "NULL argument");
    unsigned w = 0;
    CU_TRY(cudaMemcpy(&w, ex->dev.done_counter + 1, sizeof(w), cudaMemcpyDeviceToHost));
    *timed_out = w != 0;
    return RAG_OK;
}

int rag_merge_topk_dev(const double* scores_dev, const int64_t* ids_dev, int G, int B, int k, int64_t rank_stride,
                       double* out_scores_dev, int64_t* out_ids_dev, int32_t* out_counts_dev) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev || !out_counts_dev)
        return fail(RAG_EINVAL, "NULL argument");
    if (G <= 0 || B <= 0 || k <= 0 || k > RAG_MAX_K || (int64_t)G * k > 8192)
        return fail(RAG_ERANGE, "G=%d B=%d k=%d outside the supported range", G, B, k);
    if (rank_stride == 0) rank_stride = (int64_t)B * k;
    CU_TRY(merge_exact_launch(scores_dev, ids_dev, G, B, k, rank_stride, out_scores_dev, out_ids_dev, out_counts_dev,
                              g.stream));
    ++g.n_launch;
    return RAG_OK;
}

// ---------------------------------------------------------------------------
// BM25
// ---------------------------------------------------------------------------
int rag_bm25_create(rag_bm25_t** out, int64_t n_docs, int64_t n_terms, int64_t nnz, const int64_t* term_ptr,
                    const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, const double* idf,
                    double avgdl, double k1, double b) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!out || !term_ptr || !doc_len || !idf || (nnz > 0 && (!post_row || !post_tf)))
        return fail(RAG_EINVAL, "NULL argument");
    if (n_docs <= 0 || n_terms < 0 || nnz < 0 || n_docs > 0x7FFFFFF0LL) return fail(RAG_EINVAL, "bad sizes");
    if (term_ptr[0] != 0 || term_ptr[n_terms] != nnz) return fail(RAG_EINVAL, "term_ptr does not span the postings");
    rag_bm25* ix = new rag_bm25();
    ix->d.n_docs = n_docs;
    ix->d.n_terms = n_terms;
    ix->d.nnz = nnz;
    ix->h_term_ptr.assign(term_ptr, term_ptr + n_terms + 1);
    ix->h_idf.assign(idf, idf + n_terms);
    int32_t *d_tf = nullptr, *d_dl = nullptr;
    auto cleanup = [&]() {
        cudaFree(ix->d.term_ptr); cudaFree(ix->d.post_row); cudaFree(ix->d.post_impact);
        cudaFree(ix->d.idf); cudaFree(ix->d.score); cudaFree(d_tf); cudaFree(d_dl);
        delete ix;
    };
    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&ix->d.term_ptr, (size_t)(n_terms + 1) * 8);
    A((void**)&ix->d.post_row, nz * 4);
    A((void**)&ix->d.post_impact, nz * 8);
    A((void**)&ix->d.idf, (size_t)std::max<int64_t>(n_terms, 1) * 8);
    A((void**)&ix->d.score, (size_t)n_docs * 8);
    A((void**)&d_tf, nz * 4);
    A((void**)&d_dl, (size_t)n_docs * 4);
    auto C = [&](void* d, const void* h, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, g.stream);
    };
    C(ix->d.term_ptr, term_ptr, (size_t)(n_terms + 1) * 8);
    C(ix->d.post_row, post_row, (size_t)nnz * 4);
    C(d_tf, post_tf, (size_t)nnz * 4);
    C(d_dl, doc_len, (size_t)n_docs * 4);
    C(ix->d.idf, idf, (size_t)n_terms * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(ix->d.score, 0, (size_t)n_docs * 8, g.stream);
    if (e == cudaSuccess && nnz > 0) {
        e = bm25_impact_launch(ix->d.post_row, d_tf, d_dl, nnz, avgdl, k1, b, ix->d.post_impact, g.stream);
        ++g.n_launch;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) {
        cleanup();
        return fail(e == cudaErrorMemoryAllocation ? RAG_ENOMEM : RAG_ECUDA, "bm25 build: %s", cudaGetErrorString(e));
    }
    cudaFree(d_tf);
    cudaFree(d_dl);
    *out = ix;
    return RAG_OK;
}

int rag_bm25_destroy(rag_bm25_t* ix) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!ix) return RAG_OK;
    if (g.inited) {
        cudaSetDevice(g.device);
        cudaStreamSynchronize(g.stream);
        cudaFree(ix->d.term_ptr); cudaFree(ix->d.post_row); cudaFree(ix->d.post_impact);
        cudaFree(ix->d.idf); cudaFree(ix->d.score);
    }
    delete ix;
    return RAG_OK;
}

// launches the per-token accumulation of one query; fills `ranges` with the
// [lo,hi) posting ranges of its distinct scoring tokens
static int bm25_accumulate_query(rag_bm25* ix, const int32_t* terms, int nt, std::vector<int64_t>& ranges,
                                 int64_t* total) {
    ranges.clear();
    *total = 0;
    std::vector<int32_t> seen;
    for (int i = 0; i < nt; ++i) {
        const int32_t t = terms[i];
        if (t < 0 || t >= ix->d.n_terms) continue;          // (idf.get(q) or 0) == 0
        const double w = ix->h_idf[t];
        const int64_t lo = ix->h_term_ptr[t], hi = ix->h_term_ptr[t + 1];
        if (w == 0.0 || hi <= lo) continue;
        CU_TRY(bm25_accumulate_launch(ix->d, lo, hi, w, g.stream));
        ++g.n_launch;
        if (std::find(seen.begin(), seen.end(), t) == seen.end()) {
            seen.push_back(t);
            ranges.push_back(lo);
            ranges.push_back(hi);
            *total += hi - lo;
        }
    }
    return RAG_OK;
}

int rag_bm25_search(rag_bm25_t* ix, const int32_t* q_terms, const int32_t* q_ptr, int Q, int k,
                    const uint8_t* allow_bitmap, int32_t* out_rows, double* out_scores, int32_t* out_counts) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!ix || !q_ptr || !out_rows || !out_scores || !out_counts || Q <= 0) return fail(RAG_EINVAL, "NULL argument");
    if (k <= 0 || k > RAG_MAX_K) return fail(RAG_ERANGE, "k=%d outside 1..%d", k, RAG_MAX_K);
    if (q_ptr[0] != 0 || q_ptr[Q] < 0 || (q_ptr[Q] > 0 && !q_terms)) return fail(RAG_EINVAL, "bad q_ptr / q_terms");
    const int kp = std::max(16, next_pow2(k));
    for (auto& v : g.ev_valid) v = false;
    for (auto& t : g.timings) t = 0.f;
    const int n_tok = q_ptr[Q];
    const size_t tb = (size_t)std::max(n_tok, 1) * 4, pb = (size_t)(Q + 1) * 4;
    const size_t ab = allow_bitmap ? (size_t)((ix->d.n_docs + 7) / 8) : 0;
    const size_t rb = (size_t)Q * k * 4, sb = (size_t)Q * k * 8, cb = (size_t)Q * 4;
    const int n_lists = bm25_range_lists(ix->d.n_docs);
    RAG_TRY(g.bm_terms.ensure(tb));
    RAG_TRY(g.bm_ranges.ensure(pb));
    if (ab) RAG_TRY(g.bm_allow.ensure(ab + 16));
    RAG_TRY(g.bm_rows.ensure(rb));
    RAG_TRY(g.bm_scores.ensure(sb));
    RAG_TRY(g.bm_counts.ensure(cb));
    RAG_TRY(ensure_pinned(std::max(tb + pb + ab, sb + rb + cb)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(g.pinned);
    if (n_tok > 0) memcpy(pin, q_terms, (size_t)n_tok * 4);
    memcpy(pin + tb, q_ptr, pb);
    if (ab) memcpy(pin + tb + pb, allow_bitmap, ab);
    CU_TRY(cudaMemcpyAsync(g.bm_terms.p, pin, tb, cudaMemcpyHostToDevice, g.stream));
    CU_TRY(cudaMemcpyAsync(g.bm_ranges.p, pin + tb, pb, cudaMemcpyHostToDevice, g.stream));
    if (ab) CU_TRY(cudaMemcpyAsync(g.bm_allow.p, pin + tb + pb, ab, cudaMemcpyHostToDevice, g.stream));
    const uint8_t* allow_dev = ab ? g.bm_allow.as<uint8_t>() : nullptr;
    rec(0);
    const bool fast = bm25_fast_supported(ix->d.n_docs, k);
    if (fast) {
        // fast path in chunks that keep the fp64 score scratch under ~1.5 GB
        int qc = (int)std::max<int64_t>(1, std::min<int64_t>(Q, (int64_t)1500000000 / (ix->d.n_docs * 8)));
        RAG_TRY(g.bm_scratch.ensure(bm25_fast_scratch_bytes(ix->d.n_docs, k, qc)));
        for (int q0 = 0; q0 < Q; q0 += qc) {
            const int nq = std::min(qc, Q - q0);
            CU_TRY(bm25_fast_launch(ix->d, g.bm_terms.as<int32_t>(), g.bm_ranges.as<int32_t>(), q0, nq, allow_dev, k,
                                    g.bm_scratch.p, g.bm_rows.as<int32_t>(), g.bm_scores.as<double>(),
                                    g.bm_counts.as<int32_t>(), g.stream));
            g.n_launch += 4;
        }
    } else {
        RAG_TRY(g.bm_cand.ensure((size_t)Q * n_lists * kp * bm25_key_bytes()));
        CU_TRY(bm25_range_launch(ix->d, g.bm_terms.as<int32_t>(), g.bm_ranges.as<int32_t>(), nullptr, Q, allow_dev, kp, k,
                                 g.bm_cand.p, g.bm_rows.as<int32_t>(), g.bm_scores.as<double>(),
                                 g.bm_counts.as<int32_t>(), g.stream));
        g.n_launch += 2;
    }
    rec(1);
    CU_TRY(cudaMemcpyAsync(pin, g.bm_scores.p, sb, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaMemcpyAsync(pin + sb, g.bm_rows.p, rb, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaMemcpyAsync(pin + sb + rb, g.bm_counts.p, cb, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    memcpy(out_scores, pin, sb);
    memcpy(out_rows, pin + sb, rb);
    memcpy(out_counts, pin + sb + rb, cb);
    if (fast) {
        // queries whose survivor list overflowed (mass ties at the bound) are redone on the robust path
        std::vector<int32_t> redo;
        for (int q = 0; q < Q; ++q)
            if (out_counts[q] < 0) redo.push_back(q);
        if (!redo.empty()) {
            ++g.n_fallback;
            const int nr = (int)redo.size();
            RAG_TRY(g.bm_index.ensure((size_t)nr * 4));
            RAG_TRY(g.bm_cand.ensure((size_t)nr * n_lists * kp * bm25_key_bytes()));
            CU_TRY(cudaMemcpyAsync(g.bm_index.p, redo.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, g.stream));
            // outputs of the redo land in the first nr slots of the result buffers
            CU_TRY(bm25_range_launch(ix->d, g.bm_terms.as<int32_t>(), g.bm_ranges.as<int32_t>(), g.bm_index.as<int32_t>(),
                                     nr, allow_dev, kp, k, g.bm_cand.p, g.bm_rows.as<int32_t>(),
                                     g.bm_scores.as<double>(), g.bm_counts.as<int32_t>(), g.stream));
            g.n_launch += 2;
            CU_TRY(cudaMemcpyAsync(pin, g.bm_scores.p, (size_t)nr * k * 8, cudaMemcpyDeviceToHost, g.stream));
            CU_TRY(cudaMemcpyAsync(pin + sb, g.bm_rows.p, (size_t)nr * k * 4, cudaMemcpyDeviceToHost, g.stream));
            CU_TRY(cudaMemcpyAsync(pin + sb + rb, g.bm_counts.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, g.stream));
            CU_TRY(cudaStreamSynchronize(g.stream));
            for (int i = 0; i < nr; ++i) {
                const int q = redo[i];
                memcpy(out_scores + (size_t)q * k, pin + (size_t)i * k * 8, (size_t)k * 8);
                memcpy(out_rows + (size_t)q * k, pin + sb + (size_t)i * k * 4, (size_t)k * 4);
                out_counts[q] = reinterpret_cast<int32_t*>(pin + sb + rb)[i];
            }
        }
    }
    float ms = 0.f;
    if (g.ev_valid[0] && g.ev_valid[1] && cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]) == cudaSuccess) g.timings[0] = ms;
    return RAG_OK;
}

int rag_bm25_scores(rag_bm25_t* ix, const int32_t* q_terms, int n_q_terms, double* out_scores) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!ix || !out_scores || (n_q_terms > 0 && !q_terms)) return fail(RAG_EINVAL, "NULL argument");
    std::vector<int64_t> ranges;
    int64_t total = 0;
    RAG_TRY(bm25_accumulate_query(ix, q_terms, n_q_terms, ranges, &total));
    CU_TRY(cudaMemcpyAsync(out_scores, ix->d.score, (size_t)ix->d.n_docs * 8, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    const int n_ranges = (int)ranges.size() / 2;
    if (n_ranges > 0) {
        RAG_TRY(g.bm_ranges.ensure(ranges.size() * 8));
        CU_TRY(cudaMemcpyAsync(g.bm_ranges.p, ranges.data(), ranges.size() * 8, cudaMemcpyHostToDevice, g.stream));
        CU_TRY(cudaStreamSynchronize(g.stream));
        CU_TRY(bm25_reset_launch(ix->d, g.bm_ranges.as<int64_t>(), n_ranges, bm25_harvest_grid(total, g.sm_count),
                                 g.stream));
        ++g.n_launch;
        CU_TRY(cudaStreamSynchronize(g.stream));
    }
    return RAG_OK;
}

// ---------------------------------------------------------------------------
// RRF
// ---------------------------------------------------------------------------
int rag_rrf_fuse(const int32_t* ids, const double* weights, int Q, int R, int L, int rrf_k, int top, int32_t* out_ids,
                 double* out_scores, int32_t* out_counts) {
    std::lock_guard<std::mutex> lk(g.mu);
    RAG_TRY(require_init());
    if (!ids || !weights || !out_ids || !out_scores || !out_counts) return fail(RAG_EINVAL, "NULL argument");
    if (Q <= 0 || R <= 0 || L <= 0 || top <= 0) return fail(RAG_EINVAL, "sizes must be positive");
    if ((int64_t)R * L > rrf_max_entries()) return fail(RAG_ERANGE, "R*L=%lld exceeds %d", (long long)R * L,
                                                        rrf_max_entries());
    const size_t ib = (size_t)Q * R * L * 4, wb = (size_t)Q * R * 8;
    const size_t oi = (size_t)Q * top * 4, os = (size_t)Q * top * 8, oc = (size_t)Q * 4;
    RAG_TRY(g.rrf_ids.ensure(ib));
    RAG_TRY(g.rrf_w.ensure(wb));
    RAG_TRY(g.rrf_oi.ensure(oi));
    RAG_TRY(g.rrf_os.ensure(os));
    RAG_TRY(g.rrf_oc.ensure(oc));
    RAG_TRY(ensure_pinned(std::max(ib + wb, oi + os + oc)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(g.pinned);
    memcpy(pin, weights, wb);
    memcpy(pin + wb, ids, ib);
    CU_TRY(cudaMemcpyAsync(g.rrf_w.p, pin, wb, cudaMemcpyHostToDevice, g.stream));
    CU_TRY(cudaMemcpyAsync(g.rrf_ids.p, pin + wb, ib, cudaMemcpyHostToDevice, g.stream));
    CU_TRY(rrf_launch(g.rrf_ids.as<int32_t>(), g.rrf_w.as<double>(), Q, R, L, rrf_k, top, g.rrf_oi.as<int32_t>(),
                      g.rrf_os.as<double>(), g.rrf_oc.as<int32_t>(), g.stream));
    ++g.n_launch;
    CU_TRY(cudaMemcpyAsync(pin, g.rrf_os.p, os, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaMemcpyAsync(pin + os, g.rrf_oi.p, oi, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaMemcpyAsync(pin + os + oi, g.rrf_oc.p, oc, cudaMemcpyDeviceToHost, g.stream));
    CU_TRY(cudaStreamSynchronize(g.stream));
    memcpy(out_scores, pin, os);
    memcpy(out_ids, pin + os, oi);
    memcpy(out_counts, pin + os + oi, oc);
    return RAG_OK;
}

}  // extern "C"
