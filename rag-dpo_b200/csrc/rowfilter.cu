// rowfilter.cu — row predicates on the device: tombstones (deleted rows) and the compiled form of Chroma's
// `where` metadata filter, evaluated into the allowed-row bitmap the search kernels test before their top-k.
//
// Replaces the host loop over N metadata dicts (and the N/8-byte bitmap upload per call) for the filters the
// reference passes to collection.query (src/rag/pipeline.py:35-71 builds them; call sites
// src/rag/retriever.py:215-220, 380-385) and the row removal of collection.delete
// (src/processing/ingest_enterprise.py:272,304).
//
// Metadata lives on the device as one int32 code column per key (value -> code dictionaries stay in host Python;
// -1 = key missing).  A predicate is a postfix program of int32 words:
//   0 TRUE | 1 FALSE | 2 EQ col code | 3 NE col code | 4 IN col nbits nwords w0.. | 5 NIN col nbits nwords w0.. |
//   6 AND n | 7 OR n
// Chroma semantics: a missing key fails $eq / $in and passes $ne / $nin.
#include "api_common.h"
#include "common.cuh"

namespace b200rag {

__global__ void bitmap_fill_kernel(uint32_t* bm, int64_t row0, int64_t row1) {
    // one thread per 32-bit word that intersects [row0, row1)
    const int64_t w0 = row0 >> 5, w1 = (row1 + 31) >> 5;
    for (int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t m = 0xFFFFFFFFu;
        if (w == w0) m &= 0xFFFFFFFFu << (row0 & 31);
        if (w == (row1 >> 5)) m &= (1u << (row1 & 31)) - 1u;
        if (m == 0xFFFFFFFFu) bm[w] = m;
        else if (m) atomicOr(bm + w, m);
    }
}
cudaError_t bitmap_fill_launch(uint8_t* bm, int64_t row0, int64_t row1, cudaStream_t st) {
    if (row1 <= row0) return cudaSuccess;
    int64_t g = (((row1 + 31) >> 5) - (row0 >> 5) + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    bitmap_fill_kernel<<<(int)(g < 1 ? 1 : g), 256, 0, st>>>(reinterpret_cast<uint32_t*>(bm), row0, row1);
    return cudaGetLastError();
}

__global__ void bitmap_clear_rows_kernel(uint8_t* bm, const int64_t* __restrict__ rows, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rows[i];
        atomicAnd(reinterpret_cast<unsigned*>(bm) + (r >> 5), ~(1u << (r & 31)));
    }
}
cudaError_t bitmap_clear_rows_launch(uint8_t* bm, const int64_t* rows, int64_t n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int64_t g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    bitmap_clear_rows_kernel<<<(int)g, 256, 0, st>>>(bm, rows, n);
    return cudaGetLastError();
}

__global__ void bitmap_and_kernel(uint32_t* dst, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n_words) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = a[i] & b[i];
}
cudaError_t bitmap_and_launch(uint8_t* dst, const uint8_t* a, const uint8_t* b, int64_t n_rows, cudaStream_t st) {
    const int64_t n_words = (n_rows + 31) >> 5;
    if (n_words <= 0) return cudaSuccess;
    int64_t g = (n_words + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    bitmap_and_kernel<<<(int)g, 256, 0, st>>>(reinterpret_cast<uint32_t*>(dst), reinterpret_cast<const uint32_t*>(a),
                                             reinterpret_cast<const uint32_t*>(b), n_words);
    return cudaGetLastError();
}

__global__ void codes_fill_kernel(int32_t* col, int64_t row0, int64_t row1, int32_t value) {
    for (int64_t r = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < row1; r += (int64_t)gridDim.x * blockDim.x)
        col[r] = value;
}
cudaError_t codes_fill_launch(int32_t* col, int64_t row0, int64_t row1, int32_t value, cudaStream_t st) {
    if (row1 <= row0) return cudaSuccess;
    int64_t g = (row1 - row0 + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    codes_fill_kernel<<<(int)g, 256, 0, st>>>(col, row0, row1, value);
    return cudaGetLastError();
}

__device__ __forceinline__ int32_t pred_code(const PredDev& p, int col, int64_t row) {
    if (col < 0 || col >= kMaxColumns || p.cols[col] == nullptr || row >= p.col_rows[col]) return -1;
    return p.cols[col][row];
}

__device__ __forceinline__ bool pred_eval_row(const PredDev& p, int64_t row) {
    uint32_t stack = 0u;            // bit stack: bit sp-1 is the top
    int sp = 0;
    int pc = 0;
    while (pc < p.n_prog) {
        const int op = p.prog[pc];
        bool v = false;
        if (op == 0) { v = true; pc += 1; }
        else if (op == 1) { v = false; pc += 1; }
        else if (op == 2 || op == 3) {
            const int32_t c = pred_code(p, p.prog[pc + 1], row);
            v = (c == p.prog[pc + 2]) == (op == 2);
            pc += 3;
        } else if (op == 4 || op == 5) {
            const int32_t c = pred_code(p, p.prog[pc + 1], row);
            const int nbits = p.prog[pc + 2], nwords = p.prog[pc + 3];
            bool in = false;
            if (c >= 0 && c < nbits) in = ((uint32_t)p.prog[pc + 4 + (c >> 5)] >> (c & 31)) & 1u;
            v = in == (op == 4);
            pc += 4 + nwords;
        } else {                     // 6 AND n, 7 OR n: fold the n topmost entries
            const int n = p.prog[pc + 1];
            const uint32_t mask = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
            const uint32_t top = (stack >> (sp - n)) & mask;
            v = op == 6 ? (top == mask) : (top != 0u);
            sp -= n;
            stack &= sp > 0 ? ((1u << sp) - 1u) : 0u;
            pc += 2;
        }
        stack |= (v ? 1u : 0u) << sp;
        ++sp;
    }
    return sp > 0 ? ((stack >> (sp - 1)) & 1u) != 0u : true;
}

// one warp per 1024 rows: lane l evaluates rows base + 32*i + l, the ballot of round i is bitmap word base/32 + i
__global__ void __launch_bounds__(256) pred_eval_kernel(PredDev p, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_words = (p.n_rows + 31) >> 5;
    for (int64_t base = warp * 1024; base < p.n_rows; base += n_warps * 1024) {
        uint32_t mine = 0u;
        for (int i = 0; i < 32; ++i) {
            const int64_t row = base + 32 * i + lane;
            const bool ok = row < p.n_rows && pred_eval_row(p, row);
            const uint32_t w = __ballot_sync(0xffffffffu, ok);
            if (i == lane) mine = w;
        }
        const int64_t word = (base >> 5) + lane;
        if (word < n_words) {
            if (p.live) mine &= reinterpret_cast<const uint32_t*>(p.live)[word];
            out[word] = mine;
        }
    }
}

cudaError_t pred_eval_launch(const PredDev& p, uint8_t* out_bitmap, cudaStream_t st) {
    if (p.n_rows <= 0) return cudaSuccess;
    int64_t g = (p.n_rows + 1024 * 8 - 1) / (1024 * 8);
    if (g > 148 * 8) g = 148 * 8;
    pred_eval_kernel<<<(int)(g < 1 ? 1 : g), 256, 0, st>>>(p, reinterpret_cast<uint32_t*>(out_bitmap));
    return cudaGetLastError();
}

}  // namespace b200rag
