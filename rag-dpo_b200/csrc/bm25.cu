// bm25.cu — BM25 keyword scoring over CSR postings with a fused select.
//
// Replaces rank_bm25.BM25Okapi.get_scores (called at
// src/rag/bm25_index.py:153,265) and the Python select loop of
// ChunkBM25Index.search (src/rag/bm25_index.py:267-279).
//
// Bit-parity rules: fp64 throughout, numpy's evaluation order,
// no FMA contraction (explicit __dmul_rn/__ddiv_rn/__dadd_rn), and the score of
// a row is accumulated token by token in query order.  Within a token every
// posting hits a distinct row, so no atomics are needed for the accumulation.
// rag_bm25_search uses the fused bm25_range_kernel (scores of a 4096-row range in
// shared memory); rag_bm25_scores (full get_scores vector) uses one
// bm25_accumulate_kernel launch per token over a global accumulator.
//
// Algorithmic bytes per query: sum over query tokens of df(t) * (4 + 8)
// (row id + fp64 impact) + 16 * touched rows (accumulator read-modify-write).
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace b200rag {

// impact[p] = tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)), numpy order
__global__ void bm25_impact_kernel(const int32_t* __restrict__ post_row, const int32_t* __restrict__ post_tf,
                                   const int32_t* __restrict__ doc_len, int64_t nnz, double avgdl, double k1, double b,
                                   double* __restrict__ impact) {
    const double k1p1 = __dadd_rn(k1, 1.0);
    const double one_minus_b = __dsub_rn(1.0, b);
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; p < nnz; p += stride) {
        const double tf = (double)post_tf[p];
        const double dl = (double)doc_len[post_row[p]];
        const double num = __dmul_rn(tf, k1p1);
        const double ln = __ddiv_rn(__dmul_rn(b, dl), avgdl);
        const double inner = __dadd_rn(one_minus_b, ln);
        const double den = __dadd_rn(tf, __dmul_rn(k1, inner));
        impact[p] = __ddiv_rn(num, den);
    }
}

cudaError_t bm25_impact_launch(const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, int64_t nnz,
                               double avgdl, double k1, double b, double* impact, cudaStream_t st) {
    int64_t g = (nnz + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    bm25_impact_kernel<<<(int)g, 256, 0, st>>>(post_row, post_tf, doc_len, nnz, avgdl, k1, b, impact);
    return cudaGetLastError();
}

// score[row] += idf[t] * impact[p] for the postings of one token
__global__ void bm25_accumulate_kernel(const int32_t* __restrict__ post_row, const double* __restrict__ impact,
                                       int64_t lo, int64_t hi, double w, double* __restrict__ score) {
    int64_t p = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; p < hi; p += stride) {
        const int32_t r = post_row[p];
        score[r] = __dadd_rn(score[r], __dmul_rn(w, impact[p]));
    }
}

cudaError_t bm25_accumulate_launch(const Bm25Device& ix, int64_t lo, int64_t hi, double w, cudaStream_t st) {
    int64_t g = (hi - lo + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    bm25_accumulate_kernel<<<(int)g, 256, 0, st>>>(ix.post_row, ix.post_impact, lo, hi, w, ix.score);
    return cudaGetLastError();
}

// (fp64 score bits, ~row): positive doubles order like their bit patterns
struct Bm25Key {
    uint64_t s;
    uint32_t nrow;
    uint32_t pad;
    __device__ __forceinline__ bool operator<(const Bm25Key& o) const { return s < o.s || (s == o.s && nrow < o.nrow); }
    __device__ __forceinline__ bool operator>(const Bm25Key& o) const { return o < *this; }
};

int bm25_harvest_grid(int64_t total_postings, int sm_count) {
    int64_t g = (total_postings + 2047) / 2048;
    if (g > sm_count * 2) g = sm_count * 2;
    if (g < 1) g = 1;
    return (int)g;
}

// zero the accumulator on the rows a query touched (used after rag_bm25_scores,
// which reads the full vector instead of harvesting it)
__global__ void bm25_reset_kernel(const int32_t* __restrict__ post_row, const int64_t* __restrict__ ranges, int n_ranges,
                                  double* __restrict__ score) {
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int ri = 0; ri < n_ranges; ++ri)
        for (int64_t p = ranges[2 * ri] + gthread; p < ranges[2 * ri + 1]; p += nthreads) score[post_row[p]] = 0.0;
}

cudaError_t bm25_reset_launch(const Bm25Device& ix, const int64_t* d_ranges, int n_ranges, int grid, cudaStream_t st) {
    bm25_reset_kernel<<<grid, 256, 0, st>>>(ix.post_row, d_ranges, n_ranges, ix.score);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Fused per-query kernel: every CTA owns a contiguous range of kBmRange rows and
// keeps their fp64 scores in SHARED memory.  For each query token in order it adds
// the postings that fall inside its range (postings are sorted by row, so the range
// is found by binary search) with a __syncthreads() between tokens: the summation
// order per row is exactly numpy's, with no global accumulator, no atomics and no
// grid-wide barrier.  The CTA then selects its local top-kp; bm25_select_kernel
// merges the per-range lists.  grid = (row ranges, queries).
// ---------------------------------------------------------------------------
constexpr int kBmRange = 4096;        // rows per CTA: 32 KB of fp64 accumulators
constexpr int kBmMaxTokens = 128;     // query tokens handled per pass
constexpr int kBmThreads = 256;

constexpr int kBmWarps = kBmThreads / 32;
constexpr int kBmSeg = kBmRange / kBmWarps;      // rows owned by one warp: 512

__global__ void __launch_bounds__(kBmThreads)
bm25_range_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_row,
                  const double* __restrict__ impact, const double* __restrict__ idf, int64_t n_docs, int64_t n_terms,
                  const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_ptr,
                  const uint8_t* __restrict__ allow, int kp, Bm25Key* __restrict__ cand) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    double* acc = reinterpret_cast<double*>(sm_raw);                              // kBmRange
    Bm25Key* bufs = reinterpret_cast<Bm25Key*>(sm_raw + kBmRange * sizeof(double));  // 8 warps * 2 * kp
    // posting offsets of every token at the 9 warp-segment boundaries of this CTA's row range
    __shared__ int64_t s_bound[kBmMaxTokens][kBmWarps + 1];
    __shared__ double s_w[kBmMaxTokens];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * kBmRange;
    const int64_t r1 = r0 + kBmRange < n_docs ? r0 + kBmRange : n_docs;
    const int32_t* terms = q_terms + q_ptr[qi];
    const int nt = q_ptr[qi + 1] - q_ptr[qi];

    for (int i = threadIdx.x; i < kBmRange; i += kBmThreads) acc[i] = 0.0;
    for (int t0 = 0; t0 < nt; t0 += kBmMaxTokens) {
        const int tn = nt - t0 < kBmMaxTokens ? nt - t0 : kBmMaxTokens;
        __syncthreads();                             // previous pass done with s_bound / acc zeroed
        // one binary search per (token, boundary): postings of a term are sorted by row
        for (int j = threadIdx.x; j < tn * (kBmWarps + 1); j += kBmThreads) {
            const int i = j / (kBmWarps + 1), bnd = j % (kBmWarps + 1);
            const int32_t t = terms[t0 + i];
            int64_t pos = 0;
            double w = 0.0;
            if (t >= 0 && t < n_terms) {
                w = idf[t];
                int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
                int64_t target = r0 + (int64_t)bnd * kBmSeg;
                if (target > r1) target = r1;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if ((int64_t)post_row[mid] < target) lo = mid + 1; else hi = mid;
                }
                pos = lo;
            }
            s_bound[i][bnd] = pos;
            if (bnd == 0) s_w[i] = w;
        }
        __syncthreads();
        // every warp accumulates ITS 512 rows token by token (numpy's `score +=` order) with no block-wide
        // barrier; the first chunk of the next token is prefetched while the current one is added
        int64_t p_next = s_bound[0][warp] + lane;
        int32_t row_n = 0;
        double imp_n = 0.0;
        if (p_next < s_bound[0][warp + 1]) { row_n = post_row[p_next]; imp_n = impact[p_next]; }
        for (int i = 0; i < tn; ++i) {
            const double w = s_w[i];
            const int64_t hi = s_bound[i][warp + 1];
            int64_t p = p_next;
            int32_t row = row_n;
            double imp = imp_n;
            if (i + 1 < tn) {
                p_next = s_bound[i + 1][warp] + lane;
                if (p_next < s_bound[i + 1][warp + 1]) { row_n = post_row[p_next]; imp_n = impact[p_next]; }
            }
            if (w != 0.0) {
                while (p < hi) {
                    const int r = (int)(row - r0);
                    acc[r] = __dadd_rn(acc[r], __dmul_rn(w, imp));
                    p += 32;
                    if (p < hi) { row = post_row[p]; imp = impact[p]; }
                }
            }
            __syncwarp();                            // two tokens may hit the same row from different lanes
        }
    }
    // local select over this warp's rows: score > 0, allowed, (score desc, row asc)
    WarpTopKT<Bm25Key> t;
    t.init(bufs + (size_t)warp * 2 * kp, kp, lane);
    const int seg0 = warp * kBmSeg;
    for (int i0 = 0; i0 < kBmSeg; i0 += 32) {
        const int i = seg0 + i0 + lane;
        Bm25Key key{0ull, 0u, 0u};
        if (r0 + i < r1) {
            const double sc = acc[i];
            const uint32_t r = (uint32_t)(r0 + i);
            if (sc > 0.0 && bitmap_test(allow, r)) { key.s = (uint64_t)__double_as_longlong(sc); key.nrow = ~r; }
        }
        t.offer(key, lane);
    }
    t.finish(lane);
    __syncthreads();
    if (warp == 0) {                                 // fold the other warps' lists into warp 0's
        for (int w = 1; w < kBmWarps; ++w) {
            const Bm25Key* other = bufs + (size_t)w * 2 * kp;
            for (int i0 = 0; i0 < kp; i0 += 32) {
                const int i = i0 + lane;
                Bm25Key key{0ull, 0u, 0u};
                if (i < kp) key = other[i];
                t.offer(key, lane);
            }
        }
        t.finish(lane);
        Bm25Key* out = cand + ((size_t)qi * gridDim.x + blockIdx.x) * kp;
        for (int i = lane; i < kp; i += 32) out[i] = t.buf[i];
    }
}

// per-query merge of the range lists (grid = queries).  Every list is sorted descending, so the k-th
// largest list HEAD is a valid lower bound of the global k-th score (k distinct rows reach it): only list
// prefixes >= that bound can matter, which is k + a handful of entries instead of n_lists * kp.
constexpr int kBmMaxHeads = 2048;
constexpr int kBmCollect = 1024;

__global__ void __launch_bounds__(256)
bm25_select_batch_kernel(const Bm25Key* __restrict__ cand, int n_lists, int kp, int k, int32_t* out_rows,
                         double* out_scores, int32_t* out_counts) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    Bm25Key* s_keys = reinterpret_cast<Bm25Key*>(sm_raw);     // max(kBmMaxHeads, 16 * kp) entries
    __shared__ int s_n;
    __shared__ Bm25Key s_thr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x;
    const Bm25Key* src = cand + (size_t)qi * n_lists * kp;
    bool exhaustive = n_lists > kBmMaxHeads;
    int n = 0;
    if (!exhaustive) {
        int nsort = 32;
        while (nsort < n_lists) nsort <<= 1;
        for (int l = threadIdx.x; l < nsort; l += blockDim.x) {
            Bm25Key h{0ull, 0u, 0u};
            if (l < n_lists) h = src[(size_t)l * kp];
            s_keys[l] = h;
        }
        block_bitonic_desc(s_keys, nsort);
        if (threadIdx.x == 0) {
            s_thr = k <= nsort ? s_keys[k - 1] : Bm25Key{0ull, 0u, 0u};     // empty key == no bound
            s_n = 0;
        }
        __syncthreads();
        const Bm25Key thr = s_thr;
        __syncthreads();
        // walk every list while its entries reach the bound
        for (int l = threadIdx.x; l < n_lists; l += blockDim.x) {
            const Bm25Key* lp = src + (size_t)l * kp;
            for (int i = 0; i < kp; ++i) {
                const Bm25Key e = lp[i];
                if (e.s == 0ull || e < thr) break;
                const int slot = atomicAdd(&s_n, 1);
                if (slot < kBmCollect) s_keys[slot] = e;
            }
        }
        __syncthreads();
        n = s_n;
        if (n > kBmCollect) {
            exhaustive = true;                 // > 1024 entries at the bound (mass ties): take the full merge
        } else {
            int ns2 = 32;
            while (ns2 < n) ns2 <<= 1;
            for (int i = n + threadIdx.x; i < ns2; i += blockDim.x) s_keys[i] = Bm25Key{0ull, 0u, 0u};
            block_bitonic_desc(s_keys, ns2);
        }
    }
    if (exhaustive) {
        __syncthreads();
        WarpTopKT<Bm25Key> t;
        t.init(s_keys + (size_t)warp * 2 * kp, kp, lane);
        const int64_t total = (int64_t)n_lists * kp;
        const int64_t n_iter = (total + 255) / 256;
        for (int64_t it = 0; it < n_iter; ++it) {
            const int64_t i = it * 256 + threadIdx.x;
            Bm25Key key{0ull, 0u, 0u};
            if (i < total) key = src[i];
            t.offer(key, lane);
        }
        t.finish(lane);
        __syncthreads();
        block_bitonic_desc(s_keys, 8 * 2 * kp);
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        int local = 0;
        for (int i = threadIdx.x; i < k; i += blockDim.x) local += (s_keys[i].s != 0ull);
        if (local) atomicAdd(&s_n, local);
        __syncthreads();
        n = s_n;
    }
    const int nout = n < k ? n : k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < nout;
        out_rows[(size_t)qi * k + i] = ok ? (int32_t)(~s_keys[i].nrow) : -1;
        out_scores[(size_t)qi * k + i] = ok ? __longlong_as_double((long long)s_keys[i].s) : 0.0;
    }
    if (threadIdx.x == 0) out_counts[qi] = nout;
}

int bm25_range_lists(int64_t n_docs) { return (int)((n_docs + kBmRange - 1) / kBmRange); }

cudaError_t bm25_range_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr, int Q,
                              const uint8_t* allow, int kp, int k, void* cand, int32_t* out_rows, double* out_scores,
                              int32_t* out_counts, cudaStream_t st) {
    const int n_ranges = bm25_range_lists(ix.n_docs);
    size_t smem = (size_t)kBmRange * sizeof(double) + (size_t)8 * 2 * kp * sizeof(Bm25Key);
    cudaError_t e = cudaFuncSetAttribute(bm25_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (int q0 = 0; q0 < Q; q0 += 32768) {            // gridDim.y limit
        const int nq = Q - q0 < 32768 ? Q - q0 : 32768;
        dim3 grid(n_ranges, nq);
        bm25_range_kernel<<<grid, kBmThreads, smem, st>>>(ix.term_ptr, ix.post_row, ix.post_impact, ix.idf, ix.n_docs,
                                                          ix.n_terms, d_q_terms, d_q_ptr + q0, allow, kp,
                                                          reinterpret_cast<Bm25Key*>(cand) + (size_t)q0 * n_ranges * kp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const size_t entries = (size_t)std::max(kBmMaxHeads, 16 * kp);
    const size_t smem2 = entries * sizeof(Bm25Key);
    if (smem2 > 48 * 1024) {
        e = cudaFuncSetAttribute(bm25_select_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
    }
    bm25_select_batch_kernel<<<Q, 256, smem2, st>>>(reinterpret_cast<const Bm25Key*>(cand), n_ranges, kp, k, out_rows,
                                                    out_scores, out_counts);
    return cudaGetLastError();
}

size_t bm25_key_bytes() { return sizeof(Bm25Key); }

}  // namespace b200rag
