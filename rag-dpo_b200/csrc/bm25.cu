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

// first position in post_row[lo, hi) whose row is >= target (rows ascend); 8-ary: 7 independent probes per round
__device__ __forceinline__ int64_t bm25_lower_bound(const int32_t* __restrict__ post_row, int64_t lo, int64_t hi,
                                                    int64_t target) {
    while (hi - lo > 8) {
        const int64_t step = (hi - lo) >> 3;
        int32_t v[7];
#pragma unroll
        for (int u = 0; u < 7; ++u) v[u] = post_row[lo + step * (u + 1)];
        int c = 0;                            // probes below the target: a prefix (rows ascend)
#pragma unroll
        for (int u = 0; u < 7; ++u) c += (int64_t)v[u] < target;
        const int64_t base = lo;
        if (c < 7) hi = base + step * (c + 1);
        if (c > 0) lo = base + step * c + 1;
    }
    int c = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (lo + u < hi) c += (int64_t)post_row[lo + u] < target;
    return lo + c;
}

// 16-bit UPPER bound of a positive fp64 score (the top half of the fp32 rounded up, rounded up again): monotone,
// so "coarse(score) >= coarse_floor(tau)" never misses a row with score >= tau.  0 = not selectable (score <= 0).
__device__ __forceinline__ uint16_t bm25_coarse_up(double sc) {
    if (!(sc > 0.0)) return 0;
    const uint32_t b = __float_as_uint(__double2float_ru(sc));
    const uint32_t c = (b + 0xFFFFu) >> 16;
    return (uint16_t)(c > 0xFFFFu ? 0xFFFFu : c);
}
__device__ __forceinline__ uint16_t bm25_coarse_floor(double sc) {     // sc > 0
    const uint32_t c = __float_as_uint(__double2float_rd(sc)) >> 16;
    return (uint16_t)(c == 0 ? 1 : c);
}

// accumulate the scores of the CTA's row range [r0, r1) for query qi into acc (shared memory)
__device__ __forceinline__ void bm25_accumulate_range(const int64_t* __restrict__ term_ptr,
                                                      const int32_t* __restrict__ post_row,
                                                      const double* __restrict__ impact,
                                                      const double* __restrict__ idf, int64_t n_terms,
                                                      const int32_t* __restrict__ terms, int nt, int64_t r0, int64_t r1,
                                                      double* acc, int64_t (*s_bound)[kBmWarps + 1], double* s_w) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kBmRange; i += kBmThreads) acc[i] = 0.0;
    for (int t0 = 0; t0 < nt; t0 += kBmMaxTokens) {
        const int tn = nt - t0 < kBmMaxTokens ? nt - t0 : kBmMaxTokens;
        __syncthreads();                             // previous pass done with s_bound / acc zeroed
        // one search per (token, boundary): postings of a term are sorted by row.  8-ary search: 7 independent
        // probes per round, so a 500k-posting list takes 7 dependent round trips to memory instead of 19
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
                pos = bm25_lower_bound(post_row, lo, hi, target);
            }
            s_bound[i][bnd] = pos;
            if (bnd == 0) s_w[i] = w;
        }
        __syncthreads();
        // every warp accumulates ITS 512 rows token by token (numpy's `score +=` order) with no block-wide
        // barrier.  Tokens are taken in groups of kGroup: the first chunks of the whole group are requested
        // before the first one is added, and a long posting run is read 4 chunks at a time, so several loads
        // are in flight per lane instead of one
        constexpr int kGroup = 4;
        for (int i0 = 0; i0 < tn; i0 += kGroup) {
            int64_t p_g[kGroup], hi_g[kGroup];
            int32_t row_g[kGroup];
            double imp_g[kGroup], w_g[kGroup];
#pragma unroll
            for (int g = 0; g < kGroup; ++g) {
                const int i = i0 + g;
                p_g[g] = 0; hi_g[g] = 0; w_g[g] = 0.0; row_g[g] = 0; imp_g[g] = 0.0;
                if (i < tn) {
                    w_g[g] = s_w[i];
                    p_g[g] = s_bound[i][warp] + lane;
                    hi_g[g] = w_g[g] != 0.0 ? s_bound[i][warp + 1] : 0;
                    if (p_g[g] < hi_g[g]) { row_g[g] = post_row[p_g[g]]; imp_g[g] = impact[p_g[g]]; }
                }
            }
#pragma unroll
            for (int g = 0; g < kGroup; ++g) {
                if (i0 + g < tn) {
                    const double w = w_g[g];
                    const int64_t hi = hi_g[g];
                    int64_t p = p_g[g];
                    if (p < hi) {
                        const int r = (int)(row_g[g] - r0);
                        acc[r] = __dadd_rn(acc[r], __dmul_rn(w, imp_g[g]));
                        p += 32;
                    }
                    while (p < hi) {                 // within a token every posting is a different row
                        int32_t rr[4];
                        double mm[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (p + 32 * u < hi) { rr[u] = post_row[p + 32 * u]; mm[u] = impact[p + 32 * u]; }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (p + 32 * u < hi) {
                                const int r = (int)(rr[u] - r0);
                                acc[r] = __dadd_rn(acc[r], __dmul_rn(w, mm[u]));
                            }
                        p += 128;
                    }
                    __syncwarp();                    // two tokens may hit the same row from different lanes
                }
            }
        }
    }
    __syncwarp();
}

// ROBUST path: per-range sorted top-kp lists (any k, any tie structure)
__global__ void __launch_bounds__(kBmThreads)
bm25_range_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_row,
                  const double* __restrict__ impact, const double* __restrict__ idf, int64_t n_docs, int64_t n_terms,
                  const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_ptr,
                  const int32_t* __restrict__ q_index, const uint8_t* __restrict__ allow, int kp,
                  Bm25Key* __restrict__ cand) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    double* acc = reinterpret_cast<double*>(sm_raw);                              // kBmRange
    Bm25Key* bufs = reinterpret_cast<Bm25Key*>(sm_raw + kBmRange * sizeof(double));  // 8 warps * 2 * kp
    __shared__ int64_t s_bound[kBmMaxTokens][kBmWarps + 1];
    __shared__ double s_w[kBmMaxTokens];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = q_index ? q_index[blockIdx.y] : blockIdx.y;      // which query this grid row serves
    const int64_t r0 = (int64_t)blockIdx.x * kBmRange;
    const int64_t r1 = r0 + kBmRange < n_docs ? r0 + kBmRange : n_docs;
    bm25_accumulate_range(term_ptr, post_row, impact, idf, n_terms, q_terms + q_ptr[qi], q_ptr[qi + 1] - q_ptr[qi], r0,
                          r1, acc, s_bound, s_w);
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
        Bm25Key* out = cand + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kp;
        for (int i = lane; i < kp; i += 32) out[i] = t.buf[i];
    }
}

// ---------------------------------------------------------------------------
// FAST path (4 launches for a batch of queries):
//   A  bm25_scores_heads_kernel  accumulate as above, write the range's fp64 scores to global memory and
//                                its H best (score > 0, allowed) keys as "heads", H = ceil(k / n_ranges)
//   B  bm25_tau_kernel           tau = k-th largest head: k distinct rows reach it, so it is a valid lower
//                                bound of the global k-th score
//   C  bm25_filter_kernel        every row with key >= tau is appended to the query's survivor list
//   D  bm25_final_kernel         sort the (k + few) survivors; more than kBmSurvivors of them (mass ties)
//                                flags the query (count = -1) and the caller re-runs it on the robust path
// ---------------------------------------------------------------------------
constexpr int kBmSurvivors = 2048;
constexpr int kBmMaxH = 32;       // heads per range the fast path supports (static shared memory budget)

__device__ __forceinline__ Bm25Key warp_max_key(Bm25Key k) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        Bm25Key other;
        other.s = __shfl_xor_sync(0xffffffffu, (unsigned long long)k.s, o);
        other.nrow = __shfl_xor_sync(0xffffffffu, k.nrow, o);
        other.pad = 0;
        if (k < other) k = other;
    }
    return k;
}

template <bool COARSE>          // COARSE: the score vector holds 16-bit upper bounds (batched calls), else fp64
__global__ void __launch_bounds__(kBmThreads)
bm25_scores_heads_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_row,
                         const double* __restrict__ impact, const double* __restrict__ idf, int64_t n_docs,
                         int64_t n_terms, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_ptr,
                         int q0, const uint8_t* __restrict__ allow, int H, void* __restrict__ scores_out,
                         int64_t score_stride, Bm25Key* __restrict__ heads) {
    __shared__ __align__(16) double acc[kBmRange];
    __shared__ int64_t s_bound[kBmMaxTokens][kBmWarps + 1];
    __shared__ double s_w[kBmMaxTokens];
    __shared__ Bm25Key s_heads[kBmWarps][kBmMaxH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = q0 + blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * kBmRange;
    const int64_t r1 = r0 + kBmRange < n_docs ? r0 + kBmRange : n_docs;
    bm25_accumulate_range(term_ptr, post_row, impact, idf, n_terms, q_terms + q_ptr[qi], q_ptr[qi + 1] - q_ptr[qi], r0,
                          r1, acc, s_bound, s_w);
    // my 16 rows of the warp's segment: write the scores out, keep the selectable ones as keys in registers
    Bm25Key mine[kBmSeg / 32];
    double* out = reinterpret_cast<double*>(scores_out) + (size_t)blockIdx.y * score_stride;
    uint16_t* out16 = reinterpret_cast<uint16_t*>(scores_out) + (size_t)blockIdx.y * score_stride;
#pragma unroll
    for (int j = 0; j < kBmSeg / 32; ++j) {
        const int i = warp * kBmSeg + j * 32 + lane;
        Bm25Key key{0ull, 0u, 0u};
        if (r0 + i < r1) {
            const double sc = acc[i];
            if (COARSE) out16[r0 + i] = bm25_coarse_up(sc); else out[r0 + i] = sc;
            const uint32_t r = (uint32_t)(r0 + i);
            if (sc > 0.0 && bitmap_test(allow, r)) { key.s = (uint64_t)__double_as_longlong(sc); key.nrow = ~r; }
        }
        mine[j] = key;
    }
    // H rounds of warp arg-max: the warp's H best keys, descending
    for (int h = 0; h < H; ++h) {
        Bm25Key best{0ull, 0u, 0u};
#pragma unroll
        for (int j = 0; j < kBmSeg / 32; ++j)
            if (best < mine[j]) best = mine[j];
        const Bm25Key top = warp_max_key(best);
#pragma unroll
        for (int j = 0; j < kBmSeg / 32; ++j)
            if (mine[j].s == top.s && mine[j].nrow == top.nrow) mine[j] = Bm25Key{0ull, 0u, 0u};
        if (lane == 0) s_heads[warp][h] = top;
    }
    __syncthreads();
    // warp 0: the range's H best among the 8 * H warp heads
    if (warp == 0) {
        Bm25Key* dst = heads + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * H;
        const int n = kBmWarps * H;                 // <= 512
        for (int h = 0; h < H; ++h) {
            Bm25Key best{0ull, 0u, 0u};
            int where = -1;
            for (int i = lane; i < n; i += 32) {
                const Bm25Key c = s_heads[i / H][i % H];
                if (best < c) { best = c; where = i; }
            }
            const Bm25Key top = warp_max_key(best);
            if (where >= 0 && best.s == top.s && best.nrow == top.nrow) s_heads[where / H][where % H] = Bm25Key{0ull, 0u, 0u};
            __syncwarp();
            if (lane == 0) dst[h] = top;
        }
    }
}

// tau[q] = k-th largest head (empty key when fewer than k heads exist: no bound).  A block bitonic sort of the
// <= 4096 heads: 36 barrier steps for 256 heads, which beats rank counting here (245 dependent compares per thread)
__global__ void __launch_bounds__(256)
bm25_tau_kernel(const Bm25Key* __restrict__ heads, int n_heads, int k, Bm25Key* __restrict__ tau, int32_t* counts) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    Bm25Key* s_keys = reinterpret_cast<Bm25Key*>(sm_raw);
    const int q = blockIdx.x;
    int nsort = 32;
    while (nsort < n_heads) nsort <<= 1;
    for (int i = threadIdx.x; i < nsort; i += blockDim.x)
        s_keys[i] = i < n_heads ? heads[(size_t)q * n_heads + i] : Bm25Key{0ull, 0u, 0u};
    block_bitonic_desc(s_keys, nsort);
    if (threadIdx.x == 0) {
        tau[q] = k <= n_heads ? s_keys[k - 1] : Bm25Key{0ull, 0u, 0u};
        counts[q] = 0;
    }
}

template <bool COARSE>
__global__ void __launch_bounds__(256)
bm25_filter_kernel(const void* __restrict__ scores, int64_t n_docs, int64_t score_stride,
                   const uint8_t* __restrict__ allow, const Bm25Key* __restrict__ tau, Bm25Key* __restrict__ surv,
                   int32_t* __restrict__ counts) {
    const int q = blockIdx.y;
    const Bm25Key thr = tau[q];
    const double thr_s = thr.s ? __longlong_as_double((long long)thr.s) : 0.0;
    if (COARSE) {
        // 16-bit upper bounds: everything that CAN reach tau survives (a few more than k); the exact fp64 score of
        // a survivor is recomputed from the postings by bm25_final_kernel
        const uint16_t* sc = reinterpret_cast<const uint16_t*>(scores) + (size_t)q * score_stride;
        const uint16_t cthr = thr.s ? bm25_coarse_floor(thr_s) : (uint16_t)1;
        const int64_t n_vec = (n_docs + 7) >> 3;            // the stride is a multiple of 8: 16-byte vectors
        for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
            const uint4 w = *reinterpret_cast<const uint4*>(sc + 8 * v);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint16_t c = (uint16_t)(ww[j >> 1] >> (16 * (j & 1)));
                const int64_t r = 8 * v + j;
                if (c >= cthr && r < n_docs && bitmap_test(allow, (uint32_t)r)) {
                    const int slot = atomicAdd(&counts[q], 1);
                    if (slot < kBmSurvivors) surv[(size_t)q * kBmSurvivors + slot] = Bm25Key{0ull, ~(uint32_t)r, 0u};
                }
            }
        }
        return;
    }
    const double* sc = reinterpret_cast<const double*>(scores) + (size_t)q * score_stride;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_docs; r += (int64_t)gridDim.x * blockDim.x) {
        const double v = sc[r];
        if (v > 0.0 && v >= thr_s) {
            Bm25Key key{(uint64_t)__double_as_longlong(v), ~(uint32_t)r, 0u};
            if (!(key < thr) && bitmap_test(allow, (uint32_t)r)) {
                const int slot = atomicAdd(&counts[q], 1);
                if (slot < kBmSurvivors) surv[(size_t)q * kBmSurvivors + slot] = key;
            }
        }
    }
}

// exact fp64 score of one row for one query: the same products in the same (token) order as the accumulation
__device__ __forceinline__ double bm25_exact_score(const int64_t* __restrict__ term_ptr,
                                                   const int32_t* __restrict__ post_row,
                                                   const double* __restrict__ impact, const double* __restrict__ idf,
                                                   int64_t n_terms, const int32_t* __restrict__ terms, int nt,
                                                   uint32_t row) {
    double sc = 0.0;
    for (int i = 0; i < nt; ++i) {
        const int32_t t = terms[i];
        if (t < 0 || t >= n_terms) continue;
        const double w = idf[t];
        if (w == 0.0) continue;
        const int64_t hi = term_ptr[t + 1];
        const int64_t pos = bm25_lower_bound(post_row, term_ptr[t], hi, (int64_t)row);
        if (pos < hi && (uint32_t)post_row[pos] == row) sc = __dadd_rn(sc, __dmul_rn(w, impact[pos]));
    }
    return sc;
}

template <bool COARSE>
__global__ void __launch_bounds__(256)
bm25_final_kernel(const Bm25Key* __restrict__ surv, const int32_t* __restrict__ counts, int k,
                  const Bm25Key* __restrict__ tau, const int64_t* __restrict__ term_ptr,
                  const int32_t* __restrict__ post_row, const double* __restrict__ impact,
                  const double* __restrict__ idf, int64_t n_terms, const int32_t* __restrict__ q_terms,
                  const int32_t* __restrict__ q_ptr, int q0, int32_t* out_rows, double* out_scores,
                  int32_t* out_counts) {
    __shared__ Bm25Key s_keys[kBmSurvivors];
    __shared__ int s_n;
    const int q = blockIdx.x;
    int n = counts[q];
    if (n > kBmSurvivors) {                         // mass ties at the bound: robust path must redo this query
        if (threadIdx.x == 0) out_counts[q] = -1;
        return;
    }
    if (COARSE) {
        // survivors of the 16-bit filter: recompute their exact scores, keep those that really reach tau
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const Bm25Key thr = tau[q];
        const int32_t* terms = q_terms + q_ptr[q0 + q];
        const int nt = q_ptr[q0 + q + 1] - q_ptr[q0 + q];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t row = ~surv[(size_t)q * kBmSurvivors + i].nrow;
            const double sc = bm25_exact_score(term_ptr, post_row, impact, idf, n_terms, terms, nt, row);
            const Bm25Key key{(uint64_t)__double_as_longlong(sc), ~row, 0u};
            if (sc > 0.0 && !(key < thr)) s_keys[atomicAdd(&s_n, 1)] = key;
        }
        __syncthreads();
        n = s_n;
        __syncthreads();
    }
    const int nout = n < k ? n : k;
    if (n <= 256) {
        // the usual case (k + a handful of survivors): order by rank counting, no sort
        Bm25Key me{0ull, 0u, 0u};
        if ((int)threadIdx.x < n) me = COARSE ? s_keys[threadIdx.x] : surv[(size_t)q * kBmSurvivors + threadIdx.x];
        __syncthreads();
        s_keys[threadIdx.x] = me;
        __syncthreads();
        if ((int)threadIdx.x < n) {
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += me < s_keys[j];
            if (rank < k) {
                out_rows[(size_t)q * k + rank] = (int32_t)(~me.nrow);
                out_scores[(size_t)q * k + rank] = __longlong_as_double((long long)me.s);
            }
        }
        for (int i = nout + threadIdx.x; i < k; i += blockDim.x) {
            out_rows[(size_t)q * k + i] = -1;
            out_scores[(size_t)q * k + i] = 0.0;
        }
        if (threadIdx.x == 0) out_counts[q] = nout;
        return;
    }
    int nsort = 32;
    while (nsort < n) nsort <<= 1;
    for (int i = threadIdx.x; i < nsort; i += blockDim.x) {
        if (COARSE) { if (i >= n) s_keys[i] = Bm25Key{0ull, 0u, 0u}; }
        else s_keys[i] = i < n ? surv[(size_t)q * kBmSurvivors + i] : Bm25Key{0ull, 0u, 0u};
    }
    block_bitonic_desc(s_keys, nsort);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < nout;
        out_rows[(size_t)q * k + i] = ok ? (int32_t)(~s_keys[i].nrow) : -1;
        out_scores[(size_t)q * k + i] = ok ? __longlong_as_double((long long)s_keys[i].s) : 0.0;
    }
    if (threadIdx.x == 0) out_counts[q] = nout;
}

int bm25_fast_heads_per_range(int64_t n_docs, int k) {
    const int n_ranges = (int)((n_docs + kBmRange - 1) / kBmRange);
    return (k + n_ranges - 1) / n_ranges;
}
bool bm25_fast_supported(int64_t n_docs, int k) {
    const int n_ranges = (int)((n_docs + kBmRange - 1) / kBmRange);
    const int H = bm25_fast_heads_per_range(n_docs, k);
    return H <= kBmMaxH && (int64_t)n_ranges * H <= 4096;
}
size_t bm25_fast_scratch_bytes(int64_t n_docs, int k, int Q) {
    const int n_ranges = (int)((n_docs + kBmRange - 1) / kBmRange);
    const int H = bm25_fast_heads_per_range(n_docs, k);
    return (size_t)Q * (n_docs + 8) * 8 + (size_t)Q * n_ranges * H * sizeof(Bm25Key) + (size_t)Q * sizeof(Bm25Key) +
           (size_t)Q * kBmSurvivors * sizeof(Bm25Key) + (size_t)Q * 4 + 1024;
}

// queries [q0, q0+Q) of the uploaded batch; outputs written at out_* + q0 (counts = -1: redo on the robust path)
cudaError_t bm25_fast_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr, int q0, int Q,
                             const uint8_t* allow, int k, void* scratch, int32_t* out_rows, double* out_scores,
                             int32_t* out_counts, cudaStream_t st) {
    const int n_ranges = (int)((ix.n_docs + kBmRange - 1) / kBmRange);
    const int H = bm25_fast_heads_per_range(ix.n_docs, k);
    // batched calls keep 16-bit upper bounds in the score vector (a quarter of the bytes) and recompute the exact
    // score of the k + few survivors; a handful of queries keeps the fp64 vector (no recompute latency)
    const bool coarse = Q >= 8;
    const int64_t stride = (ix.n_docs + 7) & ~(int64_t)7;
    uint8_t* base = reinterpret_cast<uint8_t*>(scratch);
    void* scores = base;
    base += (size_t)Q * (ix.n_docs + 8) * 8;
    Bm25Key* heads = reinterpret_cast<Bm25Key*>(base);
    base += (size_t)Q * n_ranges * H * sizeof(Bm25Key);
    Bm25Key* tau = reinterpret_cast<Bm25Key*>(base);
    base += (size_t)Q * sizeof(Bm25Key);
    Bm25Key* surv = reinterpret_cast<Bm25Key*>(base);
    base += (size_t)Q * kBmSurvivors * sizeof(Bm25Key);
    int32_t* counts = reinterpret_cast<int32_t*>(base);
    dim3 grid_a(n_ranges, Q);
    if (coarse)
        bm25_scores_heads_kernel<true><<<grid_a, kBmThreads, 0, st>>>(ix.term_ptr, ix.post_row, ix.post_impact, ix.idf,
                                                                     ix.n_docs, ix.n_terms, d_q_terms, d_q_ptr, q0, allow,
                                                                     H, scores, stride, heads);
    else
        bm25_scores_heads_kernel<false><<<grid_a, kBmThreads, 0, st>>>(ix.term_ptr, ix.post_row, ix.post_impact, ix.idf,
                                                                      ix.n_docs, ix.n_terms, d_q_terms, d_q_ptr, q0,
                                                                      allow, H, scores, stride, heads);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n_heads = n_ranges * H;
    int nsort = 32;
    while (nsort < n_heads) nsort <<= 1;
    const size_t smem_b = (size_t)nsort * sizeof(Bm25Key);
    if (smem_b > 48 * 1024) {
        e = cudaFuncSetAttribute(bm25_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
        if (e != cudaSuccess) return e;
    }
    bm25_tau_kernel<<<Q, 256, smem_b, st>>>(heads, n_heads, k, tau, counts);
    int gx = (int)((ix.n_docs + 256 * 8 - 1) / (256 * 8));
    if (gx > 148 * 4) gx = 148 * 4;
    if (gx < 1) gx = 1;
    dim3 grid_c(gx, Q);
    int32_t* o_r = out_rows + (size_t)q0 * k;
    double* o_s = out_scores + (size_t)q0 * k;
    int32_t* o_c = out_counts + q0;
    if (coarse) {
        bm25_filter_kernel<true><<<grid_c, 256, 0, st>>>(scores, ix.n_docs, stride, allow, tau, surv, counts);
        bm25_final_kernel<true><<<Q, 256, 0, st>>>(surv, counts, k, tau, ix.term_ptr, ix.post_row, ix.post_impact, ix.idf,
                                                   ix.n_terms, d_q_terms, d_q_ptr, q0, o_r, o_s, o_c);
    } else {
        bm25_filter_kernel<false><<<grid_c, 256, 0, st>>>(scores, ix.n_docs, stride, allow, tau, surv, counts);
        bm25_final_kernel<false><<<Q, 256, 0, st>>>(surv, counts, k, tau, ix.term_ptr, ix.post_row, ix.post_impact,
                                                    ix.idf, ix.n_terms, d_q_terms, d_q_ptr, q0, o_r, o_s, o_c);
    }
    return cudaGetLastError();
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

// robust path for Q queries; q_index (device, nullable) lists which uploaded queries they are
cudaError_t bm25_range_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr,
                              const int32_t* d_q_index, int Q, const uint8_t* allow, int kp, int k, void* cand,
                              int32_t* out_rows, double* out_scores, int32_t* out_counts, cudaStream_t st) {
    const int n_ranges = bm25_range_lists(ix.n_docs);
    size_t smem = (size_t)kBmRange * sizeof(double) + (size_t)8 * 2 * kp * sizeof(Bm25Key);
    cudaError_t e = cudaFuncSetAttribute(bm25_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (int q0 = 0; q0 < Q; q0 += 32768) {            // gridDim.y limit
        const int nq = Q - q0 < 32768 ? Q - q0 : 32768;
        dim3 grid(n_ranges, nq);
        bm25_range_kernel<<<grid, kBmThreads, smem, st>>>(ix.term_ptr, ix.post_row, ix.post_impact, ix.idf, ix.n_docs,
                                                          ix.n_terms, d_q_terms, d_q_ptr + (d_q_index ? 0 : q0),
                                                          d_q_index ? d_q_index + q0 : nullptr, allow, kp,
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
