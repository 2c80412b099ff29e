// api_common.h — shared plumbing of the C ABI (include/b200rag.h): error text, device buffers, the runtime
// (devices / shard slots) and the per-handle execution context (stream, events, scratch).  Internal header.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "kernels.h"

namespace b200rag {

extern thread_local char g_err[512];
int fail(int code, const char* fmt, ...);

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

// device buffer that only ever grows; allocated on the CURRENT device (callers select it first)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need);
    // grow and keep the first `keep` bytes (stream-ordered copy on st, then synchronised)
    int grow_keep(size_t need, size_t keep, cudaStream_t st);
    void release();
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct DeviceInfo {
    int device = -1;
    int sm_count = 0, cc_major = 0, cc_minor = 0, smem_optin = 0;
};

// process-wide state: the devices this process drives.  slot i of a sharded handle runs on slots[i].
struct Runtime {
    std::mutex mu;
    bool inited = false;
    std::vector<int> slots;                 // shard slot -> CUDA device (repeats allowed: several shards per GPU)
    std::vector<DeviceInfo> devices;        // distinct devices
    cudaStream_t user_stream = nullptr;     // rag_set_stream: stream of the primary device's contexts
    std::atomic<int64_t> n_launch{0}, n_fallback{0}, n_flagged{0};
    std::mutex tmu;
    float timings[8] = {};
    int primary() const { return slots.empty() ? -1 : slots[0]; }
    const DeviceInfo* info(int device) const {
        for (auto& d : devices)
            if (d.device == device) return &d;
        return nullptr;
    }
};
extern Runtime R;
int require_init();

// execution context of one handle (or one shard of a sharded handle): its device, its own stream, events and
// page-locked staging.  Calls on different handles never share scratch, so they do not serialise on each other.
struct Ctx {
    int device = -1;
    const DeviceInfo* info = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[8] = {};
    bool ev_valid[8] = {};
    cudaEvent_t done = nullptr;             // "this shard's part of the call is queued and will be complete here"
    float timings[8] = {};
    std::recursive_mutex mu;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    int32_t* pinned_small = nullptr;        // 4 KB for flags / counters read back inside a call
    bool ready = false;

    cudaStream_t stream() const { return (R.user_stream && device == R.primary()) ? R.user_stream : own_stream; }
    int init(int device_id);
    void destroy();
    int use() const;                        // cudaSetDevice(device)
    int ensure_pinned(size_t need);
    void rec(int i) {
        if (cudaEventRecord(ev[i], stream()) == cudaSuccess) ev_valid[i] = true;
    }
    float elapsed(int a, int b) const {
        float ms = 0.f;
        if (ev_valid[a] && ev_valid[b] && cudaEventElapsedTime(&ms, ev[a], ev[b]) == cudaSuccess) return ms;
        return 0.f;
    }
    void clear_timing() {
        for (auto& v : ev_valid) v = false;
        for (auto& t : timings) t = 0.f;
    }
    void publish_timings() const {
        std::lock_guard<std::mutex> lk(R.tmu);
        for (int i = 0; i < 8; ++i) R.timings[i] = timings[i];
    }
};

// the context whose events rag_last_timings evaluates, and the host-clock entries of the last host-buffer call
extern Ctx* g_last_ctx;
extern float g_host_timings[2];
inline void set_last_ctx(Ctx* cx) {
    std::lock_guard<std::mutex> lk(R.tmu);
    g_last_ctx = cx;
}
// block-cyclic shard layout (RAG_SHARD_BLOCK rows per block)
inline int shard_of_row(int n_shards, int64_t row) {
    return n_shards == 1 ? 0 : (int)((row / RAG_SHARD_BLOCK) % n_shards);
}
inline int64_t local_of_row(int n_shards, int64_t row) {
    if (n_shards == 1) return row;
    return (row / RAG_SHARD_BLOCK / n_shards) * RAG_SHARD_BLOCK + row % RAG_SHARD_BLOCK;
}
inline int64_t shard_row_count(int n_shards, int s, int64_t n) {      // rows of shard s among the first n global rows
    if (n_shards == 1) return n;
    const int64_t nb = n / RAG_SHARD_BLOCK, rem = n % RAG_SHARD_BLOCK;
    int64_t rows = (nb / n_shards + ((nb % n_shards) > s ? 1 : 0)) * RAG_SHARD_BLOCK;
    if ((int)(nb % n_shards) == s) rows += rem;
    return rows;
}
// global bitmap -> the bytes of shard s (RAG_SHARD_BLOCK is a multiple of 8)
inline void split_bitmap_rows(int n_shards, int64_t n_total, const uint8_t* global, int s, int64_t n_local,
                              std::vector<uint8_t>& out) {
    out.assign((size_t)((n_local + 7) / 8 + 8), 0);
    const int64_t bb = RAG_SHARD_BLOCK / 8;
    const int64_t total_bytes = (n_total + 7) / 8;
    for (int64_t lb = 0; lb * RAG_SHARD_BLOCK < n_local; ++lb) {
        const int64_t gbyte = (lb * n_shards + s) * bb;
        const int64_t n = std::min<int64_t>(bb, total_bytes - gbyte);
        if (n > 0) memcpy(out.data() + lb * bb, global + gbyte, (size_t)n);
    }
}

inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

bool is_pinned_host(const void* p);

// ---- row metadata of a corpus shard / BM25 shard: tombstones + coded columns + compiled `where` predicates ----
constexpr int kMaxColumns = 32;
struct PredEntry {
    std::vector<int32_t> prog;
    uint64_t version = 0;
    DevBuf bitmap;
    uint64_t last_use = 0;
};
struct RowMeta {
    int64_t n_dead = 0;
    uint64_t version = 1;                   // bumped by every mutation of rows / tombstones / codes
    DevBuf live;                            // bitmap, bit r = row r is alive (allocated on the first delete)
    int64_t live_rows = 0;                  // rows the live bitmap covers
    DevBuf cols[kMaxColumns];               // int32 code per row, -1 = key missing
    int64_t col_rows[kMaxColumns] = {};     // rows initialised per column
    std::vector<PredEntry> cache;
    uint64_t tick = 0;
    DevBuf prog_dev, tmp_bitmap;
    void release();
};

// kernels of rowfilter.cu
cudaError_t bitmap_fill_launch(uint8_t* bm, int64_t row0, int64_t row1, cudaStream_t st);             // set bits [row0,row1)
cudaError_t bitmap_clear_rows_launch(uint8_t* bm, const int64_t* rows, int64_t n, cudaStream_t st);   // tombstones
cudaError_t bitmap_and_launch(uint8_t* dst, const uint8_t* a, const uint8_t* b, int64_t n_rows, cudaStream_t st);
cudaError_t codes_fill_launch(int32_t* col, int64_t row0, int64_t row1, int32_t value, cudaStream_t st);
struct PredDev {
    const int32_t* prog;            // device copy of the program
    int n_prog;
    const int32_t* cols[kMaxColumns];
    int64_t col_rows[kMaxColumns];
    const uint8_t* live;            // nullable
    int64_t n_rows;
};
cudaError_t pred_eval_launch(const PredDev& p, uint8_t* out_bitmap, cudaStream_t st);

}  // namespace b200rag
