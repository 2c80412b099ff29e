// api.cu — the C ABI of libb200rag.so (include/b200rag.h): runtime, corpus handles (single-GPU and sharded over
// the GPUs of one process), the orchestration of the dense kernels, the exchange step and RRF.  The BM25 entry
// points are in api_bm25.cu.  No CPU compute path exists here: without an sm_100 device every entry point fails.
#include <math.h>

#include <chrono>

#include "api_common.h"

namespace b200rag {

thread_local char g_err[512] = "";
Runtime R;
Ctx* g_last_ctx = nullptr;
float g_host_timings[2] = {};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int require_init() {
    if (!R.inited) return fail(RAG_ENODEV, "rag_init() has not succeeded: no sm_100 device bound (no CPU fallback)");
    return RAG_OK;
}

int DevBuf::ensure(size_t need) {
    if (need <= bytes) return RAG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    const size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        return fail(RAG_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    bytes = want;
    return RAG_OK;
}

int DevBuf::grow_keep(size_t need, size_t keep, cudaStream_t st) {
    if (need <= bytes) return RAG_OK;
    const size_t want = need + need / 2 + 256;
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RAG_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    if (p && keep) {
        e = cudaMemcpyAsync(np, p, std::min(keep, bytes), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(np);
            return fail(RAG_ECUDA, "growing a device buffer: %s", cudaGetErrorString(e));
        }
    }
    if (p) cudaFree(p);
    p = np;
    bytes = want;
    return RAG_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
}

void RowMeta::release() {
    live.release();
    for (auto& c : cols) c.release();
    for (auto& e : cache) e.bitmap.release();
    cache.clear();
    prog_dev.release();
    tmp_bitmap.release();
}

int Ctx::use() const {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(RAG_ENODEV, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return RAG_OK;
}

int Ctx::init(int device_id) {
    device = device_id;
    info = R.info(device_id);
    if (!info) return fail(RAG_EINVAL, "device %d was not initialised (rag_init / rag_init_devices)", device_id);
    RAG_TRY(use());
    CU_TRY(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    for (auto& e : ev) CU_TRY(cudaEventCreate(&e));
    CU_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    CU_TRY(cudaHostAlloc((void**)&pinned_small, 4096, cudaHostAllocPortable));
    memset(pinned_small, 0, 4096);
    ready = true;
    return RAG_OK;
}

void Ctx::destroy() {
    if (!ready) return;
    cudaSetDevice(device);
    cudaStreamSynchronize(stream());
    if (own_stream) cudaStreamDestroy(own_stream);
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    if (done) cudaEventDestroy(done);
    if (pinned) cudaFreeHost(pinned);
    if (pinned_small) cudaFreeHost(pinned_small);
    pinned = nullptr;
    pinned_small = nullptr;
    own_stream = nullptr;
    ready = false;
}

int Ctx::ensure_pinned(size_t need) {
    if (need <= pinned_bytes) return RAG_OK;
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_bytes = 0;
    const size_t want = need + need / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&pinned, want, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RAG_ENOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    pinned_bytes = want;
    return RAG_OK;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace b200rag

using namespace b200rag;

namespace {

int g_tc_min_batch = 2;     // smallest batch served by the tcgen05 path (env B200RAG_TC_MIN_BATCH)
int g_tc_b1_shadow = 1;     // batch-1 on fp32 corpora goes through the bf16 shadow (env B200RAG_TC_B1_SHADOW)
int g_exchange_timeout_ms = 10000;

constexpr int kDevFallbackMax = 4;       // fallback queries one stream-ordered call serves on the device
constexpr int kDevFallbackCap = 8192;    // rows per query its collect pass holds

int register_device(int device) {
    if (R.info(device)) return RAG_OK;
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RAG_ENODEV, "device %d is sm_%d%d; b200rag kernels are sm_100a only", device, prop.major, prop.minor);
    DeviceInfo d;
    d.device = device;
    d.sm_count = prop.multiProcessorCount;
    d.cc_major = prop.major;
    d.cc_minor = prop.minor;
    d.smem_optin = (int)prop.sharedMemPerBlockOptin;
    R.devices.reserve(64);                   // pointers into the vector stay valid
    R.devices.push_back(d);
    return RAG_OK;
}

int init_slots(int n, const int* devices) {
    std::lock_guard<std::mutex> lk(R.mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(RAG_ENODEV, "no CUDA device: %s (b200rag has no CPU fallback)", cudaGetErrorString(e));
    }
    if (n < 1 || n > 64 || !devices) return fail(RAG_EINVAL, "bad device list");
    for (int i = 0; i < n; ++i)
        if (devices[i] < 0 || devices[i] >= count)
            return fail(RAG_EINVAL, "device %d out of range (0..%d)", devices[i], count - 1);
    if (R.inited) {
        // idempotent for the same list; a longer list may extend a prefix (handles keep their slots)
        const size_t common = std::min<size_t>(R.slots.size(), (size_t)n);
        for (size_t i = 0; i < common; ++i)
            if (R.slots[i] != devices[i])
                return fail(RAG_EINVAL, "already bound: slot %zu is device %d (one primary device per process)", i, R.slots[i]);
        if ((size_t)n <= R.slots.size()) return RAG_OK;
    }
    for (int i = 0; i < n; ++i) RAG_TRY(register_device(devices[i]));
    // NVLink / NVSwitch peer access between every pair of distinct devices: shards write their results straight
    // into the primary device's gather buffers
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            CU_TRY(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
            if (!can) return fail(RAG_ENODEV, "device %d cannot access device %d (no peer path)", devices[i], devices[j]);
            CU_TRY(cudaSetDevice(devices[i]));
            e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(RAG_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[i], devices[j], cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    R.slots.assign(devices, devices + n);
    CU_TRY(cudaSetDevice(devices[0]));
    if (const char* v = getenv("B200RAG_TC_MIN_BATCH")) g_tc_min_batch = std::max(1, atoi(v));
    if (const char* v = getenv("B200RAG_TC_B1_SHADOW")) g_tc_b1_shadow = atoi(v) != 0;
    R.inited = true;
    return RAG_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// corpus handle
// ---------------------------------------------------------------------------
struct CorpusShard {
    Ctx cx;
    int64_t cap = 0, n = 0;
    void* rows = nullptr;
    float* max_norm = nullptr;   // device scalar
    // bf16 shadow for the tensor-core path (fp32 corpora), built lazily; appended rows extend it
    void* shadow = nullptr;
    int64_t shadow_cap = 0, shadow_rows = 0;
    float* shadow_resid = nullptr;
    RowMeta meta;
    // scratch (device)
    DevBuf q, allow, cand, cand_cnt, sample_keys, tau_keys, overflow, q16, q_resid, top, flags, tau, nflag;
    DevBuf o_rows, o_scores, o_counts, stage_f32, idx64;
    DevBuf fb_q, fb_tau, fb_counts, fb_rows, fb_scores, fb_index, fb_n, fb_done;
    int last_tc_queries = 0, last_tc_list_cap = 0;     // the tensor-core filter's candidate lists of the last call (rag_debug_last_candidates)
};

struct rag_corpus {
    int dim = 0, dtype = 0;
    size_t row_bytes = 0;
    int n_shards = 1;
    int64_t n_total = 0;
    std::vector<CorpusShard*> sh;
    std::recursive_mutex mu;
    // sharded query: gather buffers on the primary shard's device
    DevBuf g_scores, g_ids, g_counts, m_scores, m_ids, m_counts;
};

namespace {

inline int shard_of(const rag_corpus* c, int64_t row) { return shard_of_row(c->n_shards, row); }
inline int64_t local_of(const rag_corpus* c, int64_t row) { return local_of_row(c->n_shards, row); }
inline int64_t shard_rows(const rag_corpus* c, int s, int64_t n) { return shard_row_count(c->n_shards, s, n); }

int shard_reserve(rag_corpus* c, CorpusShard& sh, int64_t cap) {
    if (cap <= sh.cap) return RAG_OK;
    if (cap > 0x7FFFFFF0LL) return fail(RAG_ERANGE, "more than 2^31 rows per shard");
    RAG_TRY(sh.cx.use());
    int64_t want = std::max(cap, sh.cap + sh.cap / 2);
    void* nr = nullptr;
    cudaError_t e = cudaMalloc(&nr, (size_t)want * c->row_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = cap;
        e = cudaMalloc(&nr, (size_t)want * c->row_bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RAG_ENOMEM, "cudaMalloc(%zu) growing corpus: %s", (size_t)want * c->row_bytes, cudaGetErrorString(e));
    }
    cudaStream_t st = sh.cx.stream();
    if (sh.n > 0) {
        e = cudaMemcpyAsync(nr, sh.rows, (size_t)sh.n * c->row_bytes, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(nr);
            return fail(RAG_ECUDA, "growing corpus: %s", cudaGetErrorString(e));
        }
    } else {
        cudaStreamSynchronize(st);
    }
    cudaFree(sh.rows);
    sh.rows = nr;
    sh.cap = want;
    return RAG_OK;
}

// new rows are alive: extend the tombstone bitmap (if the shard has one) to cover rows [live_rows, n_new)
int shard_extend_live(CorpusShard& sh, int64_t n_new) {
    RowMeta& m = sh.meta;
    if (!m.live.p || n_new <= m.live_rows) return RAG_OK;
    cudaStream_t st = sh.cx.stream();
    const size_t old_bytes = (size_t)((m.live_rows + 31) / 32 * 4);
    const size_t had = m.live.bytes;
    RAG_TRY(m.live.grow_keep((size_t)((n_new + 31) / 32 * 4 + 64), old_bytes, st));
    if (m.live.bytes != had)        // a fresh allocation: everything past the copied words is undefined
        CU_TRY(cudaMemsetAsync(reinterpret_cast<uint8_t*>(m.live.p) + old_bytes, 0, m.live.bytes - old_bytes, st));
    CU_TRY(bitmap_fill_launch(m.live.as<uint8_t>(), m.live_rows, n_new, st));
    ++R.n_launch;
    CU_TRY(cudaStreamSynchronize(st));
    m.live_rows = n_new;
    return RAG_OK;
}

// rows [row0, row0+nrows) of ONE shard (local rows) from fp32 host rows
int shard_upload(rag_corpus* c, CorpusShard& sh, int64_t row0, int64_t nrows, const float* host_rows) {
    if (nrows == 0) return RAG_OK;
    RAG_TRY(sh.cx.use());
    RAG_TRY(shard_reserve(c, sh, row0 + nrows));
    cudaStream_t st = sh.cx.stream();
    const size_t chunk_rows = std::max<size_t>(1, (size_t)(32u << 20) / ((size_t)c->dim * 4));
    RAG_TRY(sh.cx.ensure_pinned(chunk_rows * c->dim * 4));
    if (c->dtype != RAG_F32) RAG_TRY(sh.stage_f32.ensure(chunk_rows * c->dim * 4));
    for (int64_t r = 0; r < nrows; r += (int64_t)chunk_rows) {
        const int64_t nr = std::min<int64_t>((int64_t)chunk_rows, nrows - r);
        const size_t bytes = (size_t)nr * c->dim * 4;
        memcpy(sh.cx.pinned, host_rows + (size_t)r * c->dim, bytes);
        uint8_t* dst = reinterpret_cast<uint8_t*>(sh.rows) + (size_t)(row0 + r) * c->row_bytes;
        if (c->dtype == RAG_F32) {
            CU_TRY(cudaMemcpyAsync(dst, sh.cx.pinned, bytes, cudaMemcpyHostToDevice, st));
        } else {
            CU_TRY(cudaMemcpyAsync(sh.stage_f32.p, sh.cx.pinned, bytes, cudaMemcpyHostToDevice, st));
            CU_TRY(convert_rows_launch(sh.stage_f32.as<float>(), dst, c->dtype, nr * c->dim, st));
            ++R.n_launch;
        }
        CU_TRY(row_norm_max_launch(dst, c->dtype, nr, c->dim, sh.max_norm, st));
        ++R.n_launch;
        CU_TRY(cudaStreamSynchronize(st));      // the pinned chunk is reused
    }
    // rows overwritten below the shadow's extent invalidate it; appended rows only extend it (incrementally)
    if (row0 < sh.shadow_rows) sh.shadow_rows = 0;
    sh.n = std::max(sh.n, row0 + nrows);
    RAG_TRY(shard_extend_live(sh, sh.n));
    ++sh.meta.version;
    return RAG_OK;
}

// splits a global row range into per-shard contiguous pieces and calls fn(shard, local_row0, nrows, offset_in_range)
template <typename F>
int for_each_piece(const rag_corpus* c, int64_t row0, int64_t nrows, F fn) {
    if (c->n_shards == 1) return nrows > 0 ? fn(0, row0, nrows, (int64_t)0) : (int)RAG_OK;
    int64_t r = row0;
    const int64_t end = row0 + nrows;
    while (r < end) {
        const int64_t blk_end = std::min(end, (r / RAG_SHARD_BLOCK + 1) * RAG_SHARD_BLOCK);
        RAG_TRY(fn(shard_of(c, r), local_of(c, r), blk_end - r, r - row0));
        r = blk_end;
    }
    return RAG_OK;
}

struct CorpusLock {
    rag_corpus* c;
    explicit CorpusLock(const rag_corpus* cc) : c(const_cast<rag_corpus*>(cc)) {
        c->mu.lock();
        for (auto* s : c->sh) s->cx.mu.lock();
    }
    ~CorpusLock() {
        for (auto it = c->sh.rbegin(); it != c->sh.rend(); ++it) (*it)->cx.mu.unlock();
        c->mu.unlock();
    }
};

void corpus_free(rag_corpus* c) {
    for (auto* s : c->sh) {
        if (!s) continue;
        if (s->cx.ready) {
            cudaSetDevice(s->cx.device);
            cudaStreamSynchronize(s->cx.stream());
        }
        cudaFree(s->rows);
        cudaFree(s->max_norm);
        cudaFree(s->shadow);
        cudaFree(s->shadow_resid);
        s->meta.release();
        for (DevBuf* b : {&s->q, &s->allow, &s->cand, &s->cand_cnt, &s->sample_keys, &s->tau_keys, &s->overflow, &s->q16,
                          &s->q_resid, &s->top, &s->flags, &s->tau, &s->nflag, &s->o_rows, &s->o_scores, &s->o_counts,
                          &s->stage_f32, &s->idx64, &s->fb_q, &s->fb_tau, &s->fb_counts, &s->fb_rows, &s->fb_scores,
                          &s->fb_index, &s->fb_n, &s->fb_done})
            b->release();
        s->cx.destroy();
        delete s;
    }
    for (DevBuf* b : {&c->g_scores, &c->g_ids, &c->g_counts, &c->m_scores, &c->m_ids, &c->m_counts}) b->release();
    cudaGetLastError();
    delete c;
}

int corpus_create(rag_corpus_t** out, int64_t capacity_rows, int dim, int dtype, int n_shards) {
    RAG_TRY(require_init());
    if (!out) return fail(RAG_EINVAL, "out is NULL");
    if (dtype != RAG_F32 && dtype != RAG_BF16 && dtype != RAG_F16) return fail(RAG_EINVAL, "bad dtype %d", dtype);
    if (dim <= 0 || dim % 64 != 0) return fail(RAG_ERANGE, "dim %d must be a positive multiple of 64", dim);
    if (dim > (dtype == RAG_F32 ? 1024 : 2048)) return fail(RAG_ERANGE, "dim %d too large for dtype %d", dim, dtype);
    if (capacity_rows < 0) return fail(RAG_EINVAL, "negative capacity");
    if (n_shards < 1 || n_shards > (int)R.slots.size())
        return fail(RAG_EINVAL, "n_shards=%d but %zu shard slot(s) are initialised (rag_init_devices)", n_shards, R.slots.size());
    if (capacity_rows > 0x7FFFFFF0LL) return fail(RAG_ERANGE, "more than 2^31 rows");
    rag_corpus* c = new rag_corpus();
    c->dim = dim;
    c->dtype = dtype;
    c->row_bytes = (size_t)dim * (dtype == RAG_F32 ? 4 : 2);
    c->n_shards = n_shards;
    for (int s = 0; s < n_shards; ++s) {
        CorpusShard* sh = new CorpusShard();
        c->sh.push_back(sh);
        int rc = sh->cx.init(R.slots[s]);
        if (rc != RAG_OK) { corpus_free(c); return rc; }
        sh->cap = std::max<int64_t>(shard_rows(c, s, capacity_rows), 1);
        cudaError_t e = cudaMalloc(&sh->rows, (size_t)sh->cap * c->row_bytes);
        if (e == cudaSuccess) e = cudaMalloc((void**)&sh->max_norm, sizeof(float));
        if (e == cudaSuccess) e = cudaMemsetAsync(sh->max_norm, 0, sizeof(float), sh->cx.stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(sh->cx.stream());
        if (e != cudaSuccess) {
            const size_t want = (size_t)sh->cap * c->row_bytes;
            cudaGetLastError();
            corpus_free(c);
            return fail(e == cudaErrorMemoryAllocation ? RAG_ENOMEM : RAG_ECUDA, "corpus shard %d (%zu bytes): %s", s, want,
                        cudaGetErrorString(e));
        }
    }
    *out = c;
    return RAG_OK;
}

}  // namespace

int bm25_set_option(const char* key, int64_t value);      // api_bm25.cu

extern "C" {

const char* rag_last_error(void) { return g_err; }
int rag_abi_version(void) { return 2; }

int rag_init(int device) { return init_slots(1, &device); }
int rag_init_devices(int n_slots, const int* devices) { return init_slots(n_slots, devices); }

int rag_slot_count(int* n_slots) {
    RAG_TRY(require_init());
    if (!n_slots) return fail(RAG_EINVAL, "NULL argument");
    *n_slots = (int)R.slots.size();
    return RAG_OK;
}

int rag_set_stream(void* cuda_stream) {
    std::lock_guard<std::mutex> lk(R.mu);
    RAG_TRY(require_init());
    R.user_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return RAG_OK;
}

int rag_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes) {
    std::lock_guard<std::mutex> lk(R.mu);
    RAG_TRY(require_init());
    const DeviceInfo* d = R.info(R.primary());
    if (sm_count) *sm_count = d->sm_count;
    if (cc_major) *cc_major = d->cc_major;
    if (cc_minor) *cc_minor = d->cc_minor;
    size_t f = 0, t = 0;
    CU_TRY(cudaSetDevice(d->device));
    CU_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return RAG_OK;
}

int rag_set_option(const char* key, int64_t value) {
    std::lock_guard<std::mutex> lk(R.mu);
    if (!key) return fail(RAG_EINVAL, "key is NULL");
    if (!strcmp(key, "tc_min_batch")) g_tc_min_batch = (int)std::max<int64_t>(1, value);
    else if (!strcmp(key, "tc_b1_shadow")) g_tc_b1_shadow = value != 0;
    else if (!strcmp(key, "sample_div")) gemm_set_sample_div((int)value);
    else if (!strcmp(key, "balance_tail")) gemm_set_balance_tail((int)value);
    else if (!strcmp(key, "pair_mode")) gemm_set_pair_mode((int)value);
    else if (!strcmp(key, "sample_resident")) gemm_set_sample_resident((int)value);
    else if (!strcmp(key, "exchange_timeout_ms")) g_exchange_timeout_ms = (int)std::max<int64_t>(1, std::min<int64_t>(value, 600000));
    else if (bm25_set_option(key, value) == RAG_OK) return RAG_OK;
    else return fail(RAG_EINVAL, "unknown option %s", key);
    return RAG_OK;
}

int rag_host_alloc(void** out, size_t bytes) {
    RAG_TRY(require_init());
    if (!out || bytes == 0) return fail(RAG_EINVAL, "bad host allocation request");
    CU_TRY(cudaSetDevice(R.primary()));
    CU_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return RAG_OK;
}

int rag_host_free(void* p) {
    if (p && R.inited) cudaFreeHost(p);
    return RAG_OK;
}

int rag_last_timings(float* ms, int n) {
    std::lock_guard<std::mutex> lk(R.tmu);
    Ctx* cx = g_last_ctx;
    if (cx && cx->ready) {
        // stream-ordered calls cannot read their events when they return: evaluate them now (complete after the
        // caller's synchronisation)
        cudaSetDevice(cx->device);
        if (cx->ev_valid[0] && cx->ev_valid[1] && cudaEventQuery(cx->ev[cx->ev_valid[3] ? 3 : 1]) == cudaSuccess) {
            R.timings[0] = cx->elapsed(0, 1);
            R.timings[1] = cx->elapsed(1, 2);
            R.timings[2] = cx->elapsed(2, 3);
            R.timings[3] = cx->ev_valid[4] && cudaEventQuery(cx->ev[4]) == cudaSuccess ? cx->elapsed(3, 4) : 0.f;
            R.timings[6] = cx->elapsed(5, 1);
        }
        cudaGetLastError();
        R.timings[4] = g_host_timings[0];
        R.timings[5] = g_host_timings[1];
    }
    for (int i = 0; i < n; ++i) ms[i] = i < 8 ? R.timings[i] : 0.f;
    return RAG_OK;
}

int rag_counters(int64_t* out, int n) {
    if (n > 0) out[0] = R.n_launch.load();
    if (n > 1) out[1] = R.n_fallback.load();
    if (n > 2) out[2] = R.n_flagged.load();
    for (int i = 3; i < n; ++i) out[i] = 0;
    return RAG_OK;
}

namespace { double eps_tc(int dim); }

// Test hook: the candidate lists the tensor-core filter of the LAST dense call on a single-shard corpus left behind
// (every row whose filter score reached the query's sample threshold): u64 keys = (order-preserving image of the
// fp32 filter score) << 32 | ~row.  out_keys: n_queries x cap_per_query, out_counts: entries written per query
// (-1: the list overflowed).  *eps_rel: the accumulation bound the margin check uses for this corpus,
// relative to |q| * max|x|.  Tests use it to measure |filter - exact| against that bound.
int rag_debug_last_candidates(const rag_corpus_t* c, int n_queries, uint64_t* out_keys, int64_t cap_per_query,
                              int32_t* out_counts, double* eps_rel) {
    RAG_TRY(require_init());
    if (!c || !out_keys || !out_counts || n_queries <= 0 || cap_per_query <= 0) return fail(RAG_EINVAL, "bad arguments");
    CorpusLock lk(c);
    if (c->n_shards != 1) return fail(RAG_EINVAL, "single-shard corpora only");
    CorpusShard& sh = *c->sh[0];
    if (sh.last_tc_queries <= 0 || n_queries > sh.last_tc_queries)
        return fail(RAG_EINVAL, "the last call did not run the tensor-core filter over %d queries", n_queries);
    RAG_TRY(sh.cx.use());
    cudaStream_t st = sh.cx.stream();
    std::vector<int32_t> cnt((size_t)n_queries);
    CU_TRY(cudaMemcpyAsync(cnt.data(), sh.cand_cnt.p, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < n_queries; ++b) {
        int64_t n = cnt[(size_t)b];
        if (n > sh.last_tc_list_cap) { out_counts[b] = -1; continue; }
        if (n > cap_per_query) n = cap_per_query;
        out_counts[b] = (int32_t)n;
        if (n > 0)
            CU_TRY(cudaMemcpyAsync(out_keys + (size_t)b * cap_per_query, sh.cand.as<uint64_t>() + (size_t)b * sh.last_tc_list_cap,
                                   (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
    if (eps_rel) *eps_rel = eps_tc(c->dim);
    return RAG_OK;
}

// ---------------------------------------------------------------------------
// corpus
// ---------------------------------------------------------------------------
int rag_corpus_create(rag_corpus_t** out, int64_t capacity_rows, int dim, int dtype) {
    return corpus_create(out, capacity_rows, dim, dtype, 1);
}
int rag_corpus_create_sharded(rag_corpus_t** out, int64_t capacity_rows, int dim, int dtype, int n_shards) {
    return corpus_create(out, capacity_rows, dim, dtype, n_shards);
}

int rag_corpus_destroy(rag_corpus_t* c) {
    if (!c) return RAG_OK;
    {
        CorpusLock lk(c);
        std::lock_guard<std::mutex> tl(R.tmu);
        for (auto* s : c->sh)
            if (g_last_ctx == &s->cx) g_last_ctx = nullptr;
    }
    corpus_free(c);
    return RAG_OK;
}

int rag_corpus_reserve(rag_corpus_t* c, int64_t capacity_rows) {
    RAG_TRY(require_init());
    if (!c) return fail(RAG_EINVAL, "corpus is NULL");
    CorpusLock lk(c);
    for (int s = 0; s < c->n_shards; ++s) RAG_TRY(shard_reserve(c, *c->sh[s], shard_rows(c, s, capacity_rows)));
    return RAG_OK;
}

int rag_corpus_upload(rag_corpus_t* c, int64_t row0, int64_t nrows, const float* host_rows) {
    RAG_TRY(require_init());
    if (!c || (!host_rows && nrows > 0)) return fail(RAG_EINVAL, "NULL argument");
    CorpusLock lk(c);
    if (row0 < 0 || nrows < 0 || row0 > c->n_total)
        return fail(RAG_EINVAL, "rows [%lld,+%lld) not contiguous with count %lld", (long long)row0, (long long)nrows,
                    (long long)c->n_total);
    RAG_TRY(for_each_piece(c, row0, nrows, [&](int s, int64_t l0, int64_t n, int64_t off) {
        return shard_upload(c, *c->sh[s], l0, n, host_rows + (size_t)off * c->dim);
    }));
    c->n_total = std::max(c->n_total, row0 + nrows);
    return RAG_OK;
}

int rag_corpus_download(const rag_corpus_t* c, int64_t row0, int64_t nrows, float* host_rows) {
    RAG_TRY(require_init());
    if (!c || (!host_rows && nrows > 0)) return fail(RAG_EINVAL, "NULL argument");
    CorpusLock lk(c);
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->n_total) return fail(RAG_EINVAL, "rows out of range");
    return for_each_piece(c, row0, nrows, [&](int s, int64_t l0, int64_t n, int64_t off) {
        CorpusShard& sh = *c->sh[s];
        RAG_TRY(sh.cx.use());
        cudaStream_t st = sh.cx.stream();
        const size_t chunk_rows = std::max<size_t>(1, (size_t)(32u << 20) / ((size_t)c->dim * 4));
        RAG_TRY(sh.cx.ensure_pinned(std::min<size_t>(chunk_rows, (size_t)n) * c->dim * 4));
        RAG_TRY(sh.stage_f32.ensure(std::min<size_t>(chunk_rows, (size_t)n) * c->dim * 4));
        for (int64_t r = 0; r < n; r += (int64_t)chunk_rows) {
            const int64_t nr = std::min<int64_t>((int64_t)chunk_rows, n - r);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(sh.rows) + (size_t)(l0 + r) * c->row_bytes;
            CU_TRY(widen_rows_launch(src, c->dtype, sh.stage_f32.as<float>(), nr * c->dim, st));
            ++R.n_launch;
            CU_TRY(cudaMemcpyAsync(sh.cx.pinned, sh.stage_f32.p, (size_t)nr * c->dim * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            memcpy(host_rows + (size_t)(off + r) * c->dim, sh.cx.pinned, (size_t)nr * c->dim * 4);
        }
        return (int)RAG_OK;
    });
}

int rag_corpus_delete_rows(rag_corpus_t* c, const int64_t* rows, int64_t n) {
    RAG_TRY(require_init());
    if (!c || (!rows && n > 0) || n < 0) return fail(RAG_EINVAL, "bad arguments");
    CorpusLock lk(c);
    std::vector<std::vector<int64_t>> per(c->n_shards);
    for (int64_t i = 0; i < n; ++i) {
        if (rows[i] < 0 || rows[i] >= c->n_total) return fail(RAG_EINVAL, "row %lld out of range", (long long)rows[i]);
        per[shard_of(c, rows[i])].push_back(local_of(c, rows[i]));
    }
    for (int s = 0; s < c->n_shards; ++s) {
        if (per[s].empty()) continue;
        CorpusShard& sh = *c->sh[s];
        RAG_TRY(sh.cx.use());
        cudaStream_t st = sh.cx.stream();
        std::sort(per[s].begin(), per[s].end());
        per[s].erase(std::unique(per[s].begin(), per[s].end()), per[s].end());
        if (!sh.meta.live.p) {                    // first delete on this shard: an all-alive bitmap
            RAG_TRY(sh.meta.live.ensure((size_t)((std::max(sh.cap, sh.n) + 31) / 32 * 4 + 64)));
            CU_TRY(cudaMemsetAsync(sh.meta.live.p, 0, sh.meta.live.bytes, st));
            CU_TRY(bitmap_fill_launch(sh.meta.live.as<uint8_t>(), 0, sh.n, st));
            ++R.n_launch;
            sh.meta.live_rows = sh.n;
        }
        // the caller (collection.delete) only lists rows that are alive: every listed row becomes a tombstone.
        // O(n) in the rows deleted: the row indices go up (8 bytes each), nothing comes back.
        const size_t nb = per[s].size();
        RAG_TRY(sh.idx64.ensure(nb * 8));
        RAG_TRY(sh.cx.ensure_pinned(nb * 8));
        memcpy(sh.cx.pinned, per[s].data(), nb * 8);
        CU_TRY(cudaMemcpyAsync(sh.idx64.p, sh.cx.pinned, nb * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(bitmap_clear_rows_launch(sh.meta.live.as<uint8_t>(), sh.idx64.as<int64_t>(), (int64_t)nb, st));
        ++R.n_launch;
        CU_TRY(cudaStreamSynchronize(st));
        sh.meta.n_dead += (int64_t)nb;
        ++sh.meta.version;
    }
    return RAG_OK;
}

int rag_corpus_compact(rag_corpus_t* c, const int64_t* keep_rows, int64_t nkeep) {
    RAG_TRY(require_init());
    if (!c) return fail(RAG_EINVAL, "corpus is NULL");
    CorpusLock lk(c);
    if (c->n_shards != 1) return fail(RAG_EINVAL, "compaction renumbers rows across shards: rebuild a sharded corpus instead");
    CorpusShard& sh = *c->sh[0];
    if ((!keep_rows && nkeep > 0) || nkeep < 0 || nkeep > sh.n) return fail(RAG_EINVAL, "bad compact arguments");
    for (int64_t i = 0; i < nkeep; ++i) {
        if (keep_rows[i] < 0 || keep_rows[i] >= sh.n || (i > 0 && keep_rows[i] <= keep_rows[i - 1]))
            return fail(RAG_EINVAL, "keep_rows must be strictly ascending rows below the count");
    }
    if (nkeep == sh.n && sh.meta.n_dead == 0) return RAG_OK;
    RAG_TRY(sh.cx.use());
    cudaStream_t st = sh.cx.stream();
    void* nr = nullptr;
    const int64_t ncap = std::max<int64_t>(nkeep, 1);
    CU_TRY(cudaMalloc(&nr, (size_t)ncap * c->row_bytes));
    if (nkeep > 0) {
        int rc = sh.idx64.ensure((size_t)nkeep * 8);
        cudaError_t e = cudaSuccess;
        if (rc == RAG_OK) {
            e = cudaMemcpyAsync(sh.idx64.p, keep_rows, (size_t)nkeep * 8, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = gather_rows_launch(sh.rows, nr, sh.idx64.as<int64_t>(), nkeep, (int)c->row_bytes, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            ++R.n_launch;
        }
        if (rc != RAG_OK || e != cudaSuccess) {       // nothing leaks on the error paths
            cudaFree(nr);
            cudaGetLastError();
            if (rc != RAG_OK) return rc;
            return fail(RAG_ECUDA, "compaction: %s", cudaGetErrorString(e));
        }
    } else {
        cudaStreamSynchronize(st);
    }
    cudaFree(sh.rows);
    sh.rows = nr;
    sh.cap = ncap;
    sh.n = nkeep;
    c->n_total = nkeep;
    sh.shadow_rows = 0;
    // tombstones are gone with the rows; the caller re-sends the coded columns of the renumbered rows
    sh.meta.live.release();
    sh.meta.live_rows = 0;
    sh.meta.n_dead = 0;
    for (int i = 0; i < kMaxColumns; ++i) { sh.meta.cols[i].release(); sh.meta.col_rows[i] = 0; }
    ++sh.meta.version;
    // the max norm only ever over-estimates after a delete, which keeps the bound valid
    return RAG_OK;
}

int rag_corpus_count(const rag_corpus_t* c, int64_t* n) {
    if (!c || !n) return fail(RAG_EINVAL, "NULL argument");
    *n = c->n_total;
    return RAG_OK;
}

int rag_corpus_live_count(const rag_corpus_t* c, int64_t* n_live) {
    if (!c || !n_live) return fail(RAG_EINVAL, "NULL argument");
    CorpusLock lk(c);
    int64_t dead = 0;
    for (auto* s : c->sh) dead += s->meta.n_dead;
    *n_live = c->n_total - dead;
    return RAG_OK;
}

int rag_corpus_set_codes(rag_corpus_t* c, int column, int64_t row0, int64_t nrows, const int32_t* codes) {
    RAG_TRY(require_init());
    if (!c || (!codes && nrows > 0)) return fail(RAG_EINVAL, "NULL argument");
    if (column < 0 || column >= kMaxColumns) return fail(RAG_ERANGE, "column %d outside 0..%d", column, kMaxColumns - 1);
    CorpusLock lk(c);
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->n_total) return fail(RAG_EINVAL, "rows out of range");
    return for_each_piece(c, row0, nrows, [&](int s, int64_t l0, int64_t n, int64_t off) {
        CorpusShard& sh = *c->sh[s];
        RAG_TRY(sh.cx.use());
        cudaStream_t st = sh.cx.stream();
        DevBuf& col = sh.meta.cols[column];
        const int64_t have = sh.meta.col_rows[column];
        const int64_t need_rows = std::max(have, l0 + n);
        RAG_TRY(col.grow_keep((size_t)std::max(need_rows, sh.cap) * 4, (size_t)have * 4, st));
        if (l0 > have) {                                // rows in between never had this key
            CU_TRY(codes_fill_launch(col.as<int32_t>(), have, l0, -1, st));
            ++R.n_launch;
        }
        RAG_TRY(sh.cx.ensure_pinned((size_t)n * 4));
        memcpy(sh.cx.pinned, codes + off, (size_t)n * 4);
        CU_TRY(cudaMemcpyAsync(col.as<int32_t>() + l0, sh.cx.pinned, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
        sh.meta.col_rows[column] = need_rows;
        ++sh.meta.version;
        return (int)RAG_OK;
    });
}

int rag_corpus_fill_synthetic(rag_corpus_t* c, uint64_t seed, int64_t gen_row0, int64_t row0, int64_t nrows) {
    RAG_TRY(require_init());
    if (!c) return fail(RAG_EINVAL, "corpus is NULL");
    CorpusLock lk(c);
    if (row0 < 0 || nrows < 0 || row0 > c->n_total) return fail(RAG_EINVAL, "rows not contiguous with count");
    RAG_TRY(for_each_piece(c, row0, nrows, [&](int s, int64_t l0, int64_t n, int64_t off) {
        CorpusShard& sh = *c->sh[s];
        RAG_TRY(sh.cx.use());
        RAG_TRY(shard_reserve(c, sh, l0 + n));
        cudaStream_t st = sh.cx.stream();
        CU_TRY(fill_synthetic_launch(sh.rows, c->dtype, l0, n, c->dim, seed, gen_row0 + off, st));
        uint8_t* dst = reinterpret_cast<uint8_t*>(sh.rows) + (size_t)l0 * c->row_bytes;
        CU_TRY(row_norm_max_launch(dst, c->dtype, n, c->dim, sh.max_norm, st));
        R.n_launch += 2;
        if (l0 < sh.shadow_rows) sh.shadow_rows = 0;
        sh.n = std::max(sh.n, l0 + n);
        ++sh.meta.version;
        return (int)RAG_OK;
    }));
    for (auto* s : c->sh) {
        RAG_TRY(s->cx.use());
        CU_TRY(cudaStreamSynchronize(s->cx.stream()));
        RAG_TRY(shard_extend_live(*s, s->n));
    }
    c->n_total = std::max(c->n_total, row0 + nrows);
    return RAG_OK;
}

int rag_corpus_device_ptr(const rag_corpus_t* c, void** rows_dev) {
    if (!c || !rows_dev) return fail(RAG_EINVAL, "NULL argument");
    if (c->n_shards != 1) return fail(RAG_EINVAL, "a sharded corpus has no single device pointer");
    *rows_dev = c->sh[0]->rows;
    return RAG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// dense top-k
// ---------------------------------------------------------------------------
namespace {

// filter error bound of the fp32 CUDA-core scan, relative to |q|*max|x|:
// <= 16 chained FMAs + 2 + 5 tree levels (every supported dim), each 2^-24 -> < 2^-19.
const double kEpsScan = 1.0 / 524288.0;
// accumulation error of the tensor-core filter relative to |q|*max|x|: dim/16 chained K=16 blocks plus the
// in-block tree, fp32 accumulators (truncating): (dim/16 + 8) * 2^-23, never below the 2^-16 the 1024-d case uses
double eps_tc(int dim) { return std::max(1.0 / 65536.0, (dim / 16 + 8) * (1.0 / 8388608.0)); }

// 16-bit operand of the tensor-core path: bf16 / fp16 rows as stored, or the bf16 shadow of fp32 rows (built
// lazily; rows appended since the last build are converted incrementally)
int ensure_bf16_operand(rag_corpus* c, CorpusShard& sh, const void** x16, const float** x_resid) {
    if (c->dtype == RAG_BF16 || c->dtype == RAG_F16) {
        *x16 = sh.rows;
        *x_resid = nullptr;
        return RAG_OK;
    }
    cudaStream_t st = sh.cx.stream();
    if (sh.shadow_rows != sh.n) {
        if (sh.shadow_cap < sh.n) {
            void* ns = nullptr;
            cudaError_t e = cudaMalloc(&ns, (size_t)sh.cap * c->dim * 2);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(RAG_ENOMEM, "cudaMalloc(%zu) for the bf16 shadow: %s", (size_t)sh.cap * c->dim * 2,
                            cudaGetErrorString(e));
            }
            if (sh.shadow && sh.shadow_rows > 0) {
                e = cudaMemcpyAsync(ns, sh.shadow, (size_t)sh.shadow_rows * c->dim * 2, cudaMemcpyDeviceToDevice, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) {
                    cudaFree(ns);
                    return fail(RAG_ECUDA, "growing the bf16 shadow: %s", cudaGetErrorString(e));
                }
            } else {
                cudaStreamSynchronize(st);
            }
            if (sh.shadow) cudaFree(sh.shadow);
            sh.shadow = ns;
            sh.shadow_cap = sh.cap;
        }
        if (!sh.shadow_resid) {
            CU_TRY(cudaMalloc((void**)&sh.shadow_resid, sizeof(float)));
            sh.shadow_rows = 0;
        }
        if (sh.shadow_rows == 0) CU_TRY(cudaMemsetAsync(sh.shadow_resid, 0, sizeof(float), st));
        // rows [shadow_rows, n): the residual bound is a running maximum (atomicMax), so appends only extend it
        const int64_t r0 = sh.shadow_rows;
        CU_TRY(shadow_launch(reinterpret_cast<const uint8_t*>(sh.rows) + (size_t)r0 * c->row_bytes, c->dtype, sh.n - r0, c->dim,
                             reinterpret_cast<uint8_t*>(sh.shadow) + (size_t)r0 * c->dim * 2, sh.shadow_resid, st));
        ++R.n_launch;
        sh.shadow_rows = sh.n;
    }
    *x16 = sh.shadow;
    *x_resid = sh.shadow_resid;
    return RAG_OK;
}

enum FallbackMode { kHostChecked = 0, kDeviceDriven = 1 };

struct DenseOut {
    int32_t* rows;       // B x k local rows (nullptr when gids is used)
    int64_t* gids;       // B x k global ids (sharded corpora), else nullptr
    double* scores;
    int32_t* counts;
};

// One slice (B <= gemm_max_batch()) of a dense call on one shard: everything is queued on the shard's stream.
//   kHostChecked:  the number of flagged queries is copied to pinned_small[0]; the caller synchronises and calls
//                  dense_finish.
//   kDeviceDriven: the exact fallback pass for up to kDevFallbackMax flagged queries is queued as well (kernels
//                  that return at once when nothing is flagged); the surplus gets counts = -1.
int dense_core_slice(rag_corpus* c, CorpusShard& sh, int shard, const float* q_dev, int B, int k, const uint8_t* allow_dev,
                     const DenseOut& out, FallbackMode mode) {
    Ctx& cx = sh.cx;
    cudaStream_t st = cx.stream();
    const DeviceInfo& di = *cx.info;
    // tensor-core path: every batch >= g_tc_min_batch (default 2); a single query only on an fp32 corpus (the
    // filter then streams the bf16 shadow: half the bytes of the fp32 rows) that is large enough to matter
    const bool use_tc = sh.n > 0 && (B >= g_tc_min_batch || (g_tc_b1_shadow && c->dtype == RAG_F32 && sh.n >= 262144));
    const int kp = use_tc ? std::max(64, next_pow2(2 * k + 1)) : std::max(16, next_pow2(k + 6));
    cx.clear_timing();
    set_last_ctx(&cx);
    RAG_TRY(sh.top.ensure((size_t)B * kp * 8));
    RAG_TRY(sh.flags.ensure((size_t)B * 4));
    RAG_TRY(sh.tau.ensure((size_t)B * 4));
    RAG_TRY(sh.nflag.ensure(4));
    CU_TRY(cudaMemsetAsync(sh.nflag.p, 0, 4, st));
    const float* q_resid = nullptr;
    const float* x_resid = nullptr;
    const uint64_t* tau_keys = nullptr;
    int32_t* overflow = nullptr;
    const int32_t* m_counts = nullptr;
    int m_flat = 0, m_lists = 0, m_len = 0;

    if (sh.n == 0) {
        CU_TRY(cudaMemsetAsync(sh.top.p, 0, (size_t)B * kp * 8, st));
    } else if (use_tc) {
        // ---- tcgen05 contraction + fused top-k (dense_gemm.cu)
        const void* x16 = nullptr;
        RAG_TRY(ensure_bf16_operand(c, sh, &x16, &x_resid));
        GemmParams p{};
        p.n_rows = sh.n;
        p.dim = c->dim;
        p.n_queries = B;
        p.kp = kp;
        p.allow = allow_dev;
        p.fp16_operands = c->dtype == RAG_F16;
        int grid = 0;
        const size_t smem = gemm_plan(p, di.sm_count, di.smem_optin, &grid);
        if (smem == 0) return fail(RAG_ERANGE, "k=%d does not fit the contraction kernel's shared memory", k);
        const int bpad = gemm_padded_queries(B);
        RAG_TRY(sh.q16.ensure((size_t)bpad * c->dim * 2));
        RAG_TRY(sh.q_resid.ensure((size_t)bpad * 4));
        const int sm = gemm_sample_m();
        RAG_TRY(sh.cand.ensure((size_t)bpad * p.list_cap * 8));
        RAG_TRY(sh.cand_cnt.ensure((size_t)bpad * 4));
        RAG_TRY(sh.sample_keys.ensure((size_t)bpad * p.n_lists * sm * 8));
        RAG_TRY(sh.tau_keys.ensure((size_t)bpad * sm * 8));
        RAG_TRY(sh.overflow.ensure((size_t)bpad * 4));
        p.cand = sh.cand.as<uint64_t>();
        p.cand_cnt = sh.cand_cnt.as<int32_t>();
        p.sample_keys = sh.sample_keys.as<uint64_t>();
        p.tau_keys = p.use_sample ? sh.tau_keys.as<uint64_t>() : nullptr;
        CU_TRY(query_prep_launch(q_dev, B, bpad, c->dim, sh.q16.p, sh.q_resid.as<float>(), p.fp16_operands, st));
        ++R.n_launch;
        q_resid = sh.q_resid.as<float>();
        CU_TRY(cudaMemsetAsync(sh.cand_cnt.p, 0, (size_t)bpad * 4, st));
        cx.rec(0);
        if (p.use_sample) {
            // sample pass -> per-query threshold (the 8th best sample score)
            CU_TRY(gemm_launch(p, 0, sh.q16.p, x16, grid, smem, st));
            CU_TRY(sample_tau_launch(sh.sample_keys.as<uint64_t>(), B, p.n_lists * sm, sm, sh.tau_keys.as<uint64_t>(), st));
            R.n_launch += 2;
        }
        cx.rec(5);                      // timings[6]: the main pass alone (the launch the roofline is quoted on)
        CU_TRY(gemm_launch(p, 1, sh.q16.p, x16, grid, smem, st));
        ++R.n_launch;
        cx.rec(1);
        sh.last_tc_queries = B;
        sh.last_tc_list_cap = p.list_cap;
        // the per-query list is one contiguous block: the 8 merge warps split it (flat count per query)
        m_counts = sh.cand_cnt.as<int32_t>();
        m_flat = 1;
        m_lists = 8;
        m_len = p.list_cap / 8;
        tau_keys = p.tau_keys;
        overflow = sh.overflow.as<int32_t>();
    } else {
        // ---- CUDA-core scan (dense_scan.cu): one launch per <= 4 queries
        ScanParams p{};
        p.rows = sh.rows;
        p.n_rows = sh.n;
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
        const size_t smem = scan_plan(p, c->dtype, di.sm_count, di.smem_optin, &grid, &nch);
        if (smem == 0) return fail(RAG_ERANGE, "k=%d does not fit the scan kernel's shared memory", k);
        const int n_lists = p.n_lists;
        RAG_TRY(sh.cand.ensure((size_t)B * n_lists * kp * 8));
        sh.last_tc_queries = 0;
        cx.rec(0);
        for (int gi = 0; gi < n_groups; ++gi) {
            ScanParams pg = p;
            pg.n_queries = std::min(4, B - gi * 4);
            pg.q = q_dev + (size_t)gi * 4 * c->dim;
            pg.cand = sh.cand.as<uint64_t>() + (size_t)gi * 4 * n_lists * kp;
            CU_TRY(scan_launch(pg, c->dtype, nch, grid, smem, st));
            ++R.n_launch;
        }
        cx.rec(1);
        m_lists = n_lists;
        m_len = kp;
    }
    cx.rec(2);
    RefineParams rp{};
    rp.top = sh.top.as<uint64_t>();
    rp.rows = sh.rows;
    rp.q = q_dev;
    rp.dtype = c->dtype;
    rp.dim = c->dim;
    rp.kp = kp;
    rp.k = k;
    rp.B = B;
    rp.eps_rel = use_tc ? eps_tc(c->dim) : kEpsScan;
    rp.q_resid = q_resid;
    rp.x_resid = x_resid;
    rp.tau_keys = tau_keys;
    rp.tau_stride = gemm_sample_m();
    rp.max_row_norm = sh.max_norm;
    rp.out_rows = out.rows;
    rp.out_gids = out.gids;
    rp.shard = shard;
    rp.n_shards = c->n_shards;
    rp.shard_block = RAG_SHARD_BLOCK;
    rp.out_scores = out.scores;
    rp.out_counts = out.counts;
    rp.flags = sh.flags.as<int32_t>();
    rp.tau = sh.tau.as<float>();
    rp.n_flagged = sh.nflag.as<int32_t>();
    const bool dev_fb = mode == kDeviceDriven && sh.n > 0;
    if (dev_fb) {
        RAG_TRY(sh.fb_q.ensure((size_t)kDevFallbackMax * c->dim * 4));
        RAG_TRY(sh.fb_tau.ensure((size_t)kDevFallbackMax * 4));
        RAG_TRY(sh.fb_counts.ensure((size_t)kDevFallbackMax * 4));
        RAG_TRY(sh.fb_index.ensure((size_t)kDevFallbackMax * 4));
        RAG_TRY(sh.fb_n.ensure(4));
        if (!sh.fb_done.p) {
            RAG_TRY(sh.fb_done.ensure(4));
            CU_TRY(cudaMemsetAsync(sh.fb_done.p, 0, 4, st));
        }
        RAG_TRY(sh.fb_rows.ensure((size_t)kDevFallbackMax * kDevFallbackCap * 4));
        RAG_TRY(sh.fb_scores.ensure((size_t)kDevFallbackMax * kDevFallbackCap * 8));
        rp.fb_done = sh.fb_done.as<unsigned>();
        rp.fb_n = sh.fb_n.as<int32_t>();
        rp.fb_index = sh.fb_index.as<int32_t>();
        rp.fb_tau = sh.fb_tau.as<float>();
        rp.fb_q = sh.fb_q.as<float>();
        rp.fb_counts = sh.fb_counts.as<unsigned>();
        rp.fb_max = kDevFallbackMax;
    }
    if (sh.n == 0) {
        CU_TRY(refine_launch(rp, st));
    } else {
        CU_TRY(merge_refine_launch(sh.cand.as<uint64_t>(), m_counts, m_flat, m_lists, m_len, use_tc ? 0 : 1, overflow, rp, st));
    }
    ++R.n_launch;
    cx.rec(3);
    if (dev_fb) {
        // ---- device-driven exact fallback: collect every row that can still reach the k-th exact score of a
        // flagged query, re-score all of them, select.  Both kernels return at once when nothing is flagged.
        ScanParams p{};
        p.rows = sh.rows;
        p.n_rows = sh.n;
        p.dim = c->dim;
        p.kp = 16;
        p.mode = 1;
        p.allow = allow_dev;
        p.n_queries = kDevFallbackMax;
        p.nq_t = scan_nq_template(kDevFallbackMax);
        p.q = sh.fb_q.as<float>();
        p.tau = sh.fb_tau.as<float>();
        p.collect_count = sh.fb_counts.as<unsigned>();
        p.collect_rows = sh.fb_rows.as<uint32_t>();
        p.collect_cap = kDevFallbackCap;
        p.n_active_dev = sh.fb_n.as<int32_t>();
        p.active_first = 0;
        int grid = 0, nch = 0;
        const size_t smem = scan_plan(p, c->dtype, di.sm_count, di.smem_optin, &grid, &nch);
        if (smem == 0) return fail(RAG_ERANGE, "fallback scan does not fit shared memory");
        CU_TRY(scan_launch(p, c->dtype, nch, grid, smem, st));
        CollectSelectParams sp{};
        sp.rows_list = sh.fb_rows.as<uint32_t>();
        sp.counts = sh.fb_counts.as<unsigned>();
        sp.cap = kDevFallbackCap;
        sp.query_index = sh.fb_index.as<int32_t>();
        sp.rows = sh.rows;
        sp.q = sh.fb_q.as<float>();
        sp.dtype = c->dtype;
        sp.dim = c->dim;
        sp.k = k;
        sp.nq = kDevFallbackMax;
        sp.scratch_scores = sh.fb_scores.as<double>();
        sp.out_rows = out.rows;
        sp.out_gids = out.gids;
        sp.shard = shard;
        sp.n_shards = c->n_shards;
        sp.shard_block = RAG_SHARD_BLOCK;
        sp.out_scores = out.scores;
        sp.out_counts = out.counts;
        sp.n_active_dev = sh.fb_n.as<int32_t>();
        CU_TRY(collect_select_launch(sp, st));
        R.n_launch += 2;
        cx.rec(4);
        // statistics only: how many queries took the pass is read when somebody asks (rag_counters stays host-side)
        return RAG_OK;
    }
    CU_TRY(cudaMemcpyAsync(cx.pinned_small, sh.nflag.p, 4, cudaMemcpyDeviceToHost, st));
    return RAG_OK;
}

// host-checked mode, after the stream was synchronised: run the exact fallback pass for the flagged queries
// (rare).  *redone is set when results were rewritten.
int dense_finish(rag_corpus* c, CorpusShard& sh, int shard, const float* q_dev, int B, int k, const uint8_t* allow_dev,
                 const DenseOut& out, bool* redone) {
    Ctx& cx = sh.cx;
    cudaStream_t st = cx.stream();
    const DeviceInfo& di = *cx.info;
    const int n_flagged = *cx.pinned_small;
    if (redone) *redone = n_flagged > 0;
    if (n_flagged <= 0) return RAG_OK;
    ++R.n_fallback;
    std::vector<int32_t> h_flags(B);
    std::vector<float> h_tau(B);
    CU_TRY(cudaMemcpyAsync(h_flags.data(), sh.flags.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h_tau.data(), sh.tau.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    std::vector<int32_t> idx;
    for (int b = 0; b < B; ++b)
        if (h_flags[b]) idx.push_back(b);
    const int nf = (int)idx.size();
    if (nf == 0) return RAG_OK;
    R.n_flagged += nf;
    RAG_TRY(sh.fb_q.ensure((size_t)std::max(nf, kDevFallbackMax) * c->dim * 4));
    RAG_TRY(sh.fb_tau.ensure((size_t)std::max(nf, kDevFallbackMax) * 4));
    RAG_TRY(sh.fb_counts.ensure((size_t)std::max(nf, kDevFallbackMax) * 4));
    RAG_TRY(sh.fb_index.ensure((size_t)std::max(nf, kDevFallbackMax) * 4));
    std::vector<float> tau_c(nf);
    for (int i = 0; i < nf; ++i) {
        tau_c[i] = h_tau[idx[i]];
        CU_TRY(cudaMemcpyAsync(sh.fb_q.as<float>() + (size_t)i * c->dim, q_dev + (size_t)idx[i] * c->dim, (size_t)c->dim * 4,
                               cudaMemcpyDeviceToDevice, st));
    }
    CU_TRY(cudaMemcpyAsync(sh.fb_tau.p, tau_c.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(sh.fb_index.p, idx.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    int cap = 4096;
    std::vector<unsigned> h_counts(nf);
    for (int attempt = 0; attempt < 2; ++attempt) {
        RAG_TRY(sh.fb_rows.ensure(std::max((size_t)nf * cap, (size_t)kDevFallbackMax * kDevFallbackCap) * 4));
        RAG_TRY(sh.fb_scores.ensure(std::max((size_t)nf * cap, (size_t)kDevFallbackMax * kDevFallbackCap) * 8));
        CU_TRY(cudaMemsetAsync(sh.fb_counts.p, 0, (size_t)nf * 4, st));
        for (int gi = 0; gi * 4 < nf; ++gi) {
            ScanParams p{};
            p.rows = sh.rows;
            p.n_rows = sh.n;
            p.dim = c->dim;
            p.kp = 16;
            p.mode = 1;
            p.allow = allow_dev;
            p.n_queries = std::min(4, nf - gi * 4);
            p.nq_t = scan_nq_template(p.n_queries);
            p.q = sh.fb_q.as<float>() + (size_t)gi * 4 * c->dim;
            p.tau = sh.fb_tau.as<float>() + gi * 4;
            p.collect_count = sh.fb_counts.as<unsigned>() + gi * 4;
            p.collect_rows = sh.fb_rows.as<uint32_t>() + (size_t)gi * 4 * cap;
            p.collect_cap = cap;
            int grid = 0, nch = 0;
            const size_t smem = scan_plan(p, c->dtype, di.sm_count, di.smem_optin, &grid, &nch);
            if (smem == 0) return fail(RAG_ERANGE, "fallback scan does not fit shared memory");
            CU_TRY(scan_launch(p, c->dtype, nch, grid, smem, st));
            ++R.n_launch;
        }
        CU_TRY(cudaMemcpyAsync(h_counts.data(), sh.fb_counts.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        unsigned mx = 0;
        for (unsigned v : h_counts) mx = std::max(mx, v);
        if (mx <= (unsigned)cap) break;
        if (attempt == 1) return fail(RAG_ECUDA, "fallback collect overflowed twice (%u > %d)", mx, cap);
        cap = (int)mx + 64;
    }
    CollectSelectParams sp{};
    sp.rows_list = sh.fb_rows.as<uint32_t>();
    sp.counts = sh.fb_counts.as<unsigned>();
    sp.cap = cap;
    sp.query_index = sh.fb_index.as<int32_t>();
    sp.rows = sh.rows;
    sp.q = sh.fb_q.as<float>();
    sp.dtype = c->dtype;
    sp.dim = c->dim;
    sp.k = k;
    sp.nq = nf;
    sp.scratch_scores = sh.fb_scores.as<double>();
    sp.out_rows = out.rows;
    sp.out_gids = out.gids;
    sp.shard = shard;
    sp.n_shards = c->n_shards;
    sp.shard_block = RAG_SHARD_BLOCK;
    sp.out_scores = out.scores;
    sp.out_counts = out.counts;
    CU_TRY(collect_select_launch(sp, st));
    ++R.n_launch;
    cx.rec(4);
    CU_TRY(cudaStreamSynchronize(st));
    return RAG_OK;
}

int dense_check(rag_corpus* c, const void* q, int B, int k, const void* o_rows, const void* o_scores,
                const void* o_counts) {
    if (!c || !q || !o_rows || !o_scores || !o_counts) return fail(RAG_EINVAL, "NULL argument");
    if (B <= 0) return fail(RAG_EINVAL, "B must be positive");
    if (k <= 0 || k > RAG_MAX_K) return fail(RAG_ERANGE, "k=%d outside 1..%d", k, RAG_MAX_K);
    return RAG_OK;
}

// the shard's allowed-row bitmap for this call: a compiled predicate (evaluated on the device, cached per
// program and corpus version), a bitmap the caller handed over (ANDed with the tombstones), the tombstones
// alone, or nothing.  host_bitmap_local must stay valid until the stream reaches the copy (pinned or synchronised).
int shard_filter(CorpusShard& sh, const int32_t* prog, int n_prog, const uint8_t* host_bitmap_local, bool bitmap_pinned,
                 const uint8_t** allow_dev) {
    cudaStream_t st = sh.cx.stream();
    RowMeta& m = sh.meta;
    const bool has_dead = m.n_dead > 0 && m.live.p;
    *allow_dev = nullptr;
    if (prog && n_prog > 0) {
        ++m.tick;
        for (auto& e : m.cache) {
            if (e.version == m.version && (int)e.prog.size() == n_prog && !memcmp(e.prog.data(), prog, (size_t)n_prog * 4)) {
                e.last_use = m.tick;
                *allow_dev = e.bitmap.as<uint8_t>();
                return RAG_OK;
            }
        }
        PredEntry* slot = nullptr;
        if (m.cache.size() < 16) {
            m.cache.emplace_back();
            slot = &m.cache.back();
        } else {
            slot = &m.cache[0];
            for (auto& e : m.cache)
                if (e.version != m.version || e.last_use < slot->last_use) slot = &e;
        }
        slot->prog.assign(prog, prog + n_prog);
        slot->version = m.version;
        slot->last_use = m.tick;
        RAG_TRY(slot->bitmap.ensure((size_t)((sh.n + 31) / 32 * 4 + 64)));
        RAG_TRY(m.prog_dev.ensure((size_t)n_prog * 4));
        CU_TRY(cudaMemcpyAsync(m.prog_dev.p, slot->prog.data(), (size_t)n_prog * 4, cudaMemcpyHostToDevice, st));
        PredDev pd{};
        pd.prog = m.prog_dev.as<int32_t>();
        pd.n_prog = n_prog;
        for (int i = 0; i < kMaxColumns; ++i) {
            pd.cols[i] = m.cols[i].as<int32_t>();
            pd.col_rows[i] = m.col_rows[i];
        }
        pd.live = has_dead ? m.live.as<uint8_t>() : nullptr;
        pd.n_rows = sh.n;
        CU_TRY(cudaMemsetAsync(slot->bitmap.p, 0, slot->bitmap.bytes, st));
        CU_TRY(pred_eval_launch(pd, slot->bitmap.as<uint8_t>(), st));
        ++R.n_launch;
        *allow_dev = slot->bitmap.as<uint8_t>();
        return RAG_OK;
    }
    if (host_bitmap_local) {
        const size_t ab = (size_t)((sh.n + 7) / 8);
        RAG_TRY(sh.allow.ensure((ab + 3) / 4 * 4 + 64));
        CU_TRY(cudaMemsetAsync(reinterpret_cast<uint8_t*>(sh.allow.p) + ab / 4 * 4, 0, 8, st));
        CU_TRY(cudaMemcpyAsync(sh.allow.p, host_bitmap_local, ab, cudaMemcpyHostToDevice, st));
        if (!bitmap_pinned) CU_TRY(cudaStreamSynchronize(st));        // pageable source: the copy must have left it
        if (has_dead) {
            CU_TRY(bitmap_and_launch(sh.allow.as<uint8_t>(), sh.allow.as<uint8_t>(), m.live.as<uint8_t>(), sh.n, st));
            ++R.n_launch;
        }
        *allow_dev = sh.allow.as<uint8_t>();
        return RAG_OK;
    }
    if (has_dead) *allow_dev = m.live.as<uint8_t>();
    return RAG_OK;
}

// host-pointer dense call on any corpus (1..G shards)
int dense_host_call(rag_corpus* c, const float* q, int B, int k, const uint8_t* allow_bitmap, const int32_t* prog,
                    int n_prog, int32_t* out_rows, double* out_scores, int32_t* out_counts) {
    const auto t_enter = std::chrono::steady_clock::now();
    const int G = c->n_shards;
    const size_t qb = (size_t)B * c->dim * 4;
    const size_t rb = (size_t)B * k * 4, sb = (size_t)B * k * 8, cb = (size_t)B * 4;
    CorpusShard& lead = *c->sh[0];
    RAG_TRY(lead.cx.use());
    const bool q_pinned = is_pinned_host(q);
    const bool out_pinned = G == 1 && is_pinned_host(out_scores) && is_pinned_host(out_rows) && is_pinned_host(out_counts);
    const size_t ab1 = (G == 1 && allow_bitmap) ? (size_t)((lead.n + 7) / 8) : 0;
    RAG_TRY(lead.cx.ensure_pinned(std::max(qb + ab1, 2 * sb + cb)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(lead.cx.pinned);
    const float* q_src = q;
    if (!q_pinned) {
        memcpy(pin, q, qb);
        q_src = reinterpret_cast<const float*>(pin);
    }
    const int step = gemm_max_batch();
    float ms_queued = 0.f;
    if (G == 1) {
        CorpusShard& sh = lead;
        cudaStream_t st = sh.cx.stream();
        RAG_TRY(sh.q.ensure(qb));
        RAG_TRY(sh.o_rows.ensure(rb));
        RAG_TRY(sh.o_scores.ensure(sb));
        RAG_TRY(sh.o_counts.ensure(cb));
        CU_TRY(cudaMemcpyAsync(sh.q.p, q_src, qb, cudaMemcpyHostToDevice, st));
        const uint8_t* bm_src = nullptr;
        if (allow_bitmap) {
            memcpy(pin + qb, allow_bitmap, ab1);
            bm_src = pin + qb;
        }
        const uint8_t* allow_dev = nullptr;
        RAG_TRY(shard_filter(sh, prog, n_prog, bm_src, true, &allow_dev));
        for (int off = 0; off < B; off += step) {
            const int nb = std::min(step, B - off);
            DenseOut o{sh.o_rows.as<int32_t>() + (size_t)off * k, nullptr, sh.o_scores.as<double>() + (size_t)off * k,
                       sh.o_counts.as<int32_t>() + off};
            RAG_TRY(dense_core_slice(c, sh, 0, sh.q.as<float>() + (size_t)off * c->dim, nb, k, allow_dev, o, kHostChecked));
            const bool last = off + nb >= B;
            for (int attempt = 0; attempt < 2; ++attempt) {
                if (last) {
                    uint8_t* ds = out_pinned ? reinterpret_cast<uint8_t*>(out_scores) : pin;
                    uint8_t* dr = out_pinned ? reinterpret_cast<uint8_t*>(out_rows) : pin + sb;
                    uint8_t* dc = out_pinned ? reinterpret_cast<uint8_t*>(out_counts) : pin + sb + rb;
                    CU_TRY(cudaMemcpyAsync(ds, sh.o_scores.p, sb, cudaMemcpyDeviceToHost, st));
                    CU_TRY(cudaMemcpyAsync(dr, sh.o_rows.p, rb, cudaMemcpyDeviceToHost, st));
                    CU_TRY(cudaMemcpyAsync(dc, sh.o_counts.p, cb, cudaMemcpyDeviceToHost, st));
                }
                if (attempt == 0 && off == 0)
                    ms_queued = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
                CU_TRY(cudaStreamSynchronize(st));
                bool redone = false;
                if (attempt == 0)
                    RAG_TRY(dense_finish(c, sh, 0, sh.q.as<float>() + (size_t)off * c->dim, nb, k, allow_dev, o, &redone));
                if (!redone || !last) break;          // otherwise copy the rewritten results once more
            }
        }
        if (!out_pinned) {
            memcpy(out_scores, pin, sb);
            memcpy(out_rows, pin + sb, rb);
            memcpy(out_counts, pin + sb + rb, cb);
        }
        sh.cx.timings[0] = 0.f;
    } else {
        // ---- one process, G shards: every shard works on its own stream and writes (score, GLOBAL id) straight
        // into the primary device's gather buffers over NVLink peer memory; the primary stream waits for the G
        // completion events, merges, and ONE synchronisation ends the call
        cudaStream_t lst = lead.cx.stream();
        RAG_TRY(c->g_scores.ensure((size_t)G * B * k * 8));
        RAG_TRY(c->g_ids.ensure((size_t)G * B * k * 8));
        RAG_TRY(c->g_counts.ensure((size_t)G * B * 4));
        RAG_TRY(c->m_scores.ensure(sb));
        RAG_TRY(c->m_ids.ensure(sb));
        RAG_TRY(c->m_counts.ensure(cb));
        std::vector<const uint8_t*> allow_devs(G, nullptr);
        std::vector<uint8_t> local_bm;
        auto shard_out = [&](int s, int off) {
            return DenseOut{nullptr, c->g_ids.as<int64_t>() + ((size_t)s * B + off) * k,
                            c->g_scores.as<double>() + ((size_t)s * B + off) * k,
                            c->g_counts.as<int32_t>() + (size_t)s * B + off};
        };
        const FallbackMode mode = B <= step ? kHostChecked : kDeviceDriven;
        for (int s = 0; s < G; ++s) {
            CorpusShard& sh = *c->sh[s];
            RAG_TRY(sh.cx.use());
            cudaStream_t st = sh.cx.stream();
            RAG_TRY(sh.q.ensure(qb));
            // pinned host memory is portable: every device DMAs the same query block over its own PCIe link
            CU_TRY(cudaMemcpyAsync(sh.q.p, q_src, qb, cudaMemcpyHostToDevice, st));
            const uint8_t* bm = nullptr;
            if (allow_bitmap) {
                split_bitmap_rows(c->n_shards, c->n_total, allow_bitmap, s, sh.n, local_bm);
                bm = local_bm.data();              // pageable: shard_filter synchronises after the copy
            }
            RAG_TRY(shard_filter(sh, prog, n_prog, bm, false, &allow_devs[s]));
            for (int off = 0; off < B; off += step) {
                const int nb = std::min(step, B - off);
                RAG_TRY(dense_core_slice(c, sh, s, sh.q.as<float>() + (size_t)off * c->dim, nb, k, allow_devs[s],
                                         shard_out(s, off), mode));
            }
            CU_TRY(cudaEventRecord(sh.cx.done, st));
        }
        RAG_TRY(lead.cx.use());
        for (int attempt = 0; attempt < 2; ++attempt) {
            for (int s = 1; s < G; ++s) CU_TRY(cudaStreamWaitEvent(lst, c->sh[s]->cx.done, 0));
            CU_TRY(merge_exact_launch(c->g_scores.as<double>(), c->g_ids.as<int64_t>(), G, B, k, (int64_t)B * k,
                                      c->m_scores.as<double>(), c->m_ids.as<int64_t>(), c->m_counts.as<int32_t>(), lst));
            ++R.n_launch;
            CU_TRY(cudaMemcpyAsync(pin, c->m_scores.p, sb, cudaMemcpyDeviceToHost, lst));
            CU_TRY(cudaMemcpyAsync(pin + sb, c->m_ids.p, sb, cudaMemcpyDeviceToHost, lst));
            CU_TRY(cudaMemcpyAsync(pin + 2 * sb, c->m_counts.p, cb, cudaMemcpyDeviceToHost, lst));
            if (attempt == 0)
                ms_queued = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
            CU_TRY(cudaStreamSynchronize(lst));
            bool any = false;
            if (attempt == 0 && mode == kHostChecked) {
                for (int s = 0; s < G; ++s) {
                    CorpusShard& sh = *c->sh[s];
                    RAG_TRY(sh.cx.use());
                    bool redone = false;
                    RAG_TRY(dense_finish(c, sh, s, sh.q.as<float>(), B, k, allow_devs[s], shard_out(s, 0), &redone));
                    if (redone) {
                        CU_TRY(cudaEventRecord(sh.cx.done, sh.cx.stream()));
                        any = true;
                    }
                }
                RAG_TRY(lead.cx.use());
            }
            if (!any) break;
        }
        // a shard may have hit a surplus of deep ties in a device-driven slice (counts = -1 in ITS list): the
        // merged list would silently miss rows.  Check the shards' counts (G x B ints).
        if (mode == kDeviceDriven) {
            std::vector<int32_t> gc((size_t)G * B);
            CU_TRY(cudaMemcpy(gc.data(), c->g_counts.p, gc.size() * 4, cudaMemcpyDeviceToHost));
            for (int32_t v : gc)
                if (v < 0) return fail(RAG_ERANGE, "a query needs the exact fallback pass: call in batches of <= %d", step);
        }
        memcpy(out_scores, pin, sb);
        const int64_t* ids = reinterpret_cast<const int64_t*>(pin + sb);
        for (size_t i = 0; i < (size_t)B * k; ++i) out_rows[i] = (int32_t)ids[i];
        memcpy(out_counts, pin + 2 * sb, cb);
    }
    g_host_timings[0] = ms_queued;
    g_host_timings[1] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return RAG_OK;
}

}  // namespace

extern "C" {

int rag_dense_topk_dev(rag_corpus_t* c, const float* q_dev, int B, int k, const uint8_t* allow_bitmap_dev,
                       int32_t* out_rows_dev, double* out_scores_dev, int32_t* out_counts_dev) {
    RAG_TRY(require_init());
    RAG_TRY(dense_check(c, q_dev, B, k, out_rows_dev, out_scores_dev, out_counts_dev));
    CorpusLock lk(c);
    if (c->n_shards != 1) return fail(RAG_EINVAL, "device-pointer calls serve single-shard corpora");
    CorpusShard& sh = *c->sh[0];
    RAG_TRY(sh.cx.use());
    const uint8_t* allow = allow_bitmap_dev;
    if (!allow && sh.meta.n_dead > 0 && sh.meta.live.p) allow = sh.meta.live.as<uint8_t>();
    const int step = gemm_max_batch();
    for (int off = 0; off < B; off += step) {
        const int nb = std::min(step, B - off);
        DenseOut o{out_rows_dev + (size_t)off * k, nullptr, out_scores_dev + (size_t)off * k, out_counts_dev + off};
        RAG_TRY(dense_core_slice(c, sh, 0, q_dev + (size_t)off * c->dim, nb, k, allow, o, kDeviceDriven));
    }
    g_host_timings[0] = g_host_timings[1] = 0.f;
    return RAG_OK;
}

int rag_dense_topk(rag_corpus_t* c, const float* q, int B, int k, const uint8_t* allow_bitmap, int32_t* out_rows,
                   double* out_scores, int32_t* out_counts) {
    RAG_TRY(require_init());
    RAG_TRY(dense_check(c, q, B, k, out_rows, out_scores, out_counts));
    CorpusLock lk(c);
    return dense_host_call(c, q, B, k, allow_bitmap, nullptr, 0, out_rows, out_scores, out_counts);
}

int rag_dense_topk_where(rag_corpus_t* c, const float* q, int B, int k, const int32_t* where_prog, int n_words,
                         int32_t* out_rows, double* out_scores, int32_t* out_counts) {
    RAG_TRY(require_init());
    RAG_TRY(dense_check(c, q, B, k, out_rows, out_scores, out_counts));
    if (n_words < 0 || n_words > (1 << 20) || (n_words > 0 && !where_prog)) return fail(RAG_EINVAL, "bad where program");
    CorpusLock lk(c);
    return dense_host_call(c, q, B, k, nullptr, where_prog, n_words, out_rows, out_scores, out_counts);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// peer-memory exchange (exchange.cu), merge, RRF
// ---------------------------------------------------------------------------
struct rag_exchange {
    Ctx* cx = nullptr;                        // stream bookkeeping (the primary device)
    int world = 0, rank = 0;
    size_t slot_bytes = 0, total_bytes = 0;
    uint8_t* local = nullptr;                 // payload[2][world][slot] | flags[2][world] | counter | error word
    void* mapped[kMaxExchangeRanks] = {};     // peers' buffers as opened in this process (mine: == local)
    bool connected = false;
    uint64_t epoch = 0;
    cudaStream_t stream = nullptr;            // every step runs on the stream of the first one
    ExchangeDev dev{};
    std::mutex mu;
};

namespace {

Ctx* g_misc_ctx = nullptr;       // primary-device context for handle-less calls (merge, RRF, exchange)
DevBuf g_rrf_ids, g_rrf_w, g_rrf_oi, g_rrf_os, g_rrf_oc;

int misc_ctx(Ctx** out) {
    std::lock_guard<std::mutex> lk(R.mu);
    if (!g_misc_ctx) {
        Ctx* cx = new Ctx();
        int rc = cx->init(R.primary());
        if (rc != RAG_OK) {
            delete cx;
            return rc;
        }
        g_misc_ctx = cx;
    }
    *out = g_misc_ctx;
    return RAG_OK;
}

size_t exchange_flags_offset(const rag_exchange* ex) { return 2 * (size_t)ex->world * ex->slot_bytes; }

int exchange_step(rag_exchange_t* ex, const double* my_scores_dev, const int64_t* my_ids_dev, const int32_t* my_rows_dev,
                  int64_t row_lo, const int32_t* my_counts_dev, int B, int k, double* out_scores_dev, int64_t* out_ids_dev,
                  int32_t* out_counts_dev) {
    RAG_TRY(require_init());
    if (!ex || !my_scores_dev || (!my_ids_dev && !my_rows_dev) || !out_scores_dev || !out_ids_dev || !out_counts_dev)
        return fail(RAG_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ex->mu);
    if (!ex->connected) return fail(RAG_EINVAL, "exchange not connected");
    if (B <= 0 || k <= 0 || k > RAG_MAX_K || (int64_t)ex->world * k > 8192)
        return fail(RAG_ERANGE, "world=%d B=%d k=%d outside the supported range", ex->world, B, k);
    if ((size_t)B * k * 16 > ex->slot_bytes)
        return fail(RAG_ERANGE, "B*k*16 = %zu bytes exceed the exchange slot (%zu)", (size_t)B * k * 16, ex->slot_bytes);
    RAG_TRY(ex->cx->use());
    cudaStream_t st = ex->cx->stream();
    // the two buffer parities are only safe when every step of this exchange is ordered on ONE stream
    if (ex->epoch == 0) ex->stream = st;
    else if (ex->stream != st)
        return fail(RAG_EINVAL, "the exchange is pinned to the stream of its first step (rag_set_stream changed it)");
    ex->dev.timeout_cycles = (long long)g_exchange_timeout_ms * 2000000LL;
    // the epoch advances only once the push is queued: a failed launch must not desynchronise the ranks
    CU_TRY(exchange_push_launch(ex->dev, my_scores_dev, my_ids_dev, my_rows_dev, row_lo, my_counts_dev, B, k,
                                ex->epoch + 1, st));
    ++ex->epoch;
    ++R.n_launch;
    CU_TRY(exchange_merge_launch(ex->dev, B, k, ex->epoch, out_scores_dev, out_ids_dev, out_counts_dev, st));
    ++R.n_launch;
    return RAG_OK;
}

}  // namespace

extern "C" {

int rag_exchange_create(rag_exchange_t** out, int world, int rank, size_t slot_bytes, void* handle_out) {
    RAG_TRY(require_init());
    if (!out || !handle_out) return fail(RAG_EINVAL, "NULL argument");
    if (world < 1 || world > kMaxExchangeRanks || rank < 0 || rank >= world || slot_bytes == 0)
        return fail(RAG_ERANGE, "world=%d rank=%d slot_bytes=%zu outside the supported range", world, rank, slot_bytes);
    static_assert(sizeof(cudaIpcMemHandle_t) == RAG_IPC_HANDLE_BYTES, "handle size");
    Ctx* cx = nullptr;
    RAG_TRY(misc_ctx(&cx));
    RAG_TRY(cx->use());
    rag_exchange* ex = new rag_exchange();
    ex->cx = cx;
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
    RAG_TRY(require_init());
    if (!ex || !handles) return fail(RAG_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ex->mu);
    if (ex->connected) return fail(RAG_EINVAL, "exchange already connected");
    RAG_TRY(ex->cx->use());
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
                    if (q != ex->rank && ex->mapped[q]) {
                        cudaIpcCloseMemHandle(ex->mapped[q]);
                        ex->mapped[q] = nullptr;
                    }
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
    if (!ex) return RAG_OK;
    if (R.inited && ex->cx) {
        cudaSetDevice(ex->cx->device);
        cudaDeviceSynchronize();
        for (int r = 0; r < ex->world; ++r)
            if (r != ex->rank && ex->mapped[r]) cudaIpcCloseMemHandle(ex->mapped[r]);
        if (ex->local) cudaFree(ex->local);
        cudaGetLastError();
    }
    delete ex;
    return RAG_OK;
}

int rag_exchange_merge_topk_dev(rag_exchange_t* ex, const double* my_scores_dev, const int64_t* my_ids_dev, int B, int k,
                                double* out_scores_dev, int64_t* out_ids_dev, int32_t* out_counts_dev) {
    return exchange_step(ex, my_scores_dev, my_ids_dev, nullptr, 0, nullptr, B, k, out_scores_dev, out_ids_dev, out_counts_dev);
}

int rag_exchange_merge_rows_dev(rag_exchange_t* ex, const double* my_scores_dev, const int32_t* my_rows_dev,
                                int64_t row_lo, const int32_t* my_counts_dev, int B, int k, double* out_scores_dev,
                                int64_t* out_ids_dev, int32_t* out_counts_dev) {
    return exchange_step(ex, my_scores_dev, nullptr, my_rows_dev, row_lo, my_counts_dev, B, k, out_scores_dev, out_ids_dev,
                         out_counts_dev);
}

int rag_exchange_status(rag_exchange_t* ex, int* timed_out) {
    RAG_TRY(require_init());
    if (!ex || !timed_out) return fail(RAG_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(ex->mu);
    if (!ex->connected) return fail(RAG_EINVAL, "exchange not connected");
    RAG_TRY(ex->cx->use());
    unsigned w = 0;
    CU_TRY(cudaMemcpy(&w, ex->dev.done_counter + 1, sizeof(w), cudaMemcpyDeviceToHost));
    *timed_out = w != 0;
    return RAG_OK;
}

int rag_merge_topk_dev(const double* scores_dev, const int64_t* ids_dev, int G, int B, int k, int64_t rank_stride,
                       double* out_scores_dev, int64_t* out_ids_dev, int32_t* out_counts_dev) {
    RAG_TRY(require_init());
    if (!scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev || !out_counts_dev)
        return fail(RAG_EINVAL, "NULL argument");
    if (G <= 0 || B <= 0 || k <= 0 || k > RAG_MAX_K || (int64_t)G * k > 8192)
        return fail(RAG_ERANGE, "G=%d B=%d k=%d outside the supported range", G, B, k);
    if (rank_stride == 0) rank_stride = (int64_t)B * k;
    Ctx* cx = nullptr;
    RAG_TRY(misc_ctx(&cx));
    std::lock_guard<std::recursive_mutex> lk(cx->mu);
    RAG_TRY(cx->use());
    CU_TRY(merge_exact_launch(scores_dev, ids_dev, G, B, k, rank_stride, out_scores_dev, out_ids_dev, out_counts_dev,
                              cx->stream()));
    ++R.n_launch;
    return RAG_OK;
}

int rag_rrf_fuse(const int32_t* ids, const double* weights, int Q, int R_, int L, int rrf_k, int top, int32_t* out_ids,
                 double* out_scores, int32_t* out_counts) {
    RAG_TRY(require_init());
    if (!ids || !weights || !out_ids || !out_scores || !out_counts) return fail(RAG_EINVAL, "NULL argument");
    if (Q <= 0 || R_ <= 0 || L <= 0 || top <= 0) return fail(RAG_EINVAL, "sizes must be positive");
    if ((int64_t)R_ * L > rrf_max_entries())
        return fail(RAG_ERANGE, "R*L=%lld exceeds %d", (long long)R_ * L, rrf_max_entries());
    Ctx* cx = nullptr;
    RAG_TRY(misc_ctx(&cx));
    std::lock_guard<std::recursive_mutex> lk(cx->mu);
    RAG_TRY(cx->use());
    cudaStream_t st = cx->stream();
    const size_t ib = (size_t)Q * R_ * L * 4, wb = (size_t)Q * R_ * 8;
    const size_t oi = (size_t)Q * top * 4, os = (size_t)Q * top * 8, oc = (size_t)Q * 4;
    RAG_TRY(g_rrf_ids.ensure(ib));
    RAG_TRY(g_rrf_w.ensure(wb));
    RAG_TRY(g_rrf_oi.ensure(oi));
    RAG_TRY(g_rrf_os.ensure(os));
    RAG_TRY(g_rrf_oc.ensure(oc));
    RAG_TRY(cx->ensure_pinned(std::max(ib + wb, oi + os + oc)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(cx->pinned);
    memcpy(pin, weights, wb);
    memcpy(pin + wb, ids, ib);
    CU_TRY(cudaMemcpyAsync(g_rrf_w.p, pin, wb, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(g_rrf_ids.p, pin + wb, ib, cudaMemcpyHostToDevice, st));
    CU_TRY(rrf_launch(g_rrf_ids.as<int32_t>(), g_rrf_w.as<double>(), Q, R_, L, rrf_k, top, g_rrf_oi.as<int32_t>(),
                      g_rrf_os.as<double>(), g_rrf_oc.as<int32_t>(), st));
    ++R.n_launch;
    CU_TRY(cudaMemcpyAsync(pin, g_rrf_os.p, os, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pin + os, g_rrf_oi.p, oi, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pin + os + oi, g_rrf_oc.p, oc, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    memcpy(out_scores, pin, os);
    memcpy(out_ids, pin + os, oi);
    memcpy(out_counts, pin + os + oi, oc);
    return RAG_OK;
}

int rag_rerank_select(const float* scores, const double* boosts, const int32_t* lens, int Q, int L, int top_k,
                      double min_score, int32_t* out_idx, double* out_scores, int32_t* out_counts) {
    RAG_TRY(require_init());
    if (!scores || !out_idx || !out_scores || !out_counts) return fail(RAG_EINVAL, "NULL argument");
    if (Q <= 0 || L <= 0) return fail(RAG_EINVAL, "sizes must be positive");
    if (top_k < 3) return fail(RAG_ERANGE, "top_k=%d: the step always returns at least 3 candidates", top_k);
    if (L > rerank_max_candidates()) return fail(RAG_ERANGE, "L=%d exceeds %d", L, rerank_max_candidates());
    Ctx* cx = nullptr;
    RAG_TRY(misc_ctx(&cx));
    std::lock_guard<std::recursive_mutex> lk(cx->mu);
    RAG_TRY(cx->use());
    cudaStream_t st = cx->stream();
    const size_t sb = (size_t)Q * L * 4, bb = boosts ? (size_t)Q * L * 8 : 0, lb = lens ? (size_t)Q * 4 : 0;
    const size_t oi = (size_t)Q * top_k * 4, os = (size_t)Q * top_k * 8, oc = (size_t)Q * 4;
    RAG_TRY(g_rrf_w.ensure(bb + 8));                   // boosts (8-byte aligned first)
    RAG_TRY(g_rrf_ids.ensure(sb + lb));                // scores | lens
    RAG_TRY(g_rrf_oi.ensure(oi));
    RAG_TRY(g_rrf_os.ensure(os));
    RAG_TRY(g_rrf_oc.ensure(oc));
    RAG_TRY(cx->ensure_pinned(std::max(bb + sb + lb, oi + os + oc)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(cx->pinned);
    if (bb) memcpy(pin, boosts, bb);
    memcpy(pin + bb, scores, sb);
    if (lb) memcpy(pin + bb + sb, lens, lb);
    if (bb) CU_TRY(cudaMemcpyAsync(g_rrf_w.p, pin, bb, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(g_rrf_ids.p, pin + bb, sb + lb, cudaMemcpyHostToDevice, st));
    CU_TRY(rerank_select_launch(g_rrf_ids.as<float>(), bb ? g_rrf_w.as<double>() : nullptr,
                                lb ? reinterpret_cast<const int32_t*>(reinterpret_cast<uint8_t*>(g_rrf_ids.p) + sb) : nullptr, Q, L,
                                top_k, min_score, g_rrf_oi.as<int32_t>(), g_rrf_os.as<double>(), g_rrf_oc.as<int32_t>(), st));
    ++R.n_launch;
    CU_TRY(cudaMemcpyAsync(pin, g_rrf_os.p, os, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pin + os, g_rrf_oi.p, oi, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pin + os + oi, g_rrf_oc.p, oc, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    memcpy(out_scores, pin, os);
    memcpy(out_idx, pin + os, oi);
    memcpy(out_counts, pin + os + oi, oc);
    return RAG_OK;
}

}  // extern "C"
