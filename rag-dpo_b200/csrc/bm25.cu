// bm25.cu — BM25 keyword scoring over CSR postings with a fused select.
//
// Replaces rank_bm25.BM25Okapi.get_scores (called at
// src/rag/bm25_index.py:153,265) and the Python select loop of
// ChunkBM25Index.search (src/rag/bm25_index.py:267-279).
//
// Bit-parity rules: the scores that are RETURNED are fp64 throughout, numpy's evaluation order, no
// FMA contraction (explicit __dmul_rn/__ddiv_rn/__dadd_rn), accumulated token by token in query order.
//   rag_bm25_search  fast path: an integer FILTER over the packed postings / dense columns (fixed-point upper bounds of
//                    idf*impact; bm25_resolve_kernel + bm25_filter_tma_kernel / bm25_filter_kernel) finds k + a handful of survivors per
//                    query, whose exact fp64 scores are recomputed from the postings (bm25_finish_kernel);
//                    robust path: fp64 accumulators of a 4096-row range in shared memory (bm25_range_kernel);
//   rag_bm25_scores  (full get_scores vector) one bm25_accumulate_kernel launch per token over a global accumulator.
//
// Algorithmic bytes per query of the fast path: sum over the query's tokens of the term's list in the format that
// serves it — 4 x df (packed stream) or 2 x n_docs (dense column); exact path: df x 12 (row id + fp64 impact).
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
// FAST path (3 launches for a batch of queries).
//
// The product idf[t] * impact[p] of a posting does not depend on the query, so the index holds it a second time
// as an 18-bit fixed-point UPPER bound q (unit = max product / ~2^18, rounded up, +1):
//   * packed with the 14-bit local row of its 16384-row BLOCK, 4 bytes per posting (every term), and
//   * for DENSE terms (df >= n_docs / 8) as a 16-bit COLUMN over all rows: ceil(q / 4), 0 where the term does not
//     occur — 2 bytes per row, read with fully coalesced loads and added into REGISTERS (a thread owns fixed rows
//     of the tile), no row ids, no shared-memory traffic, no search.
// The filter pass adds these integers:
//     U[row] = sum over the query's tokens of q   (resp. 4 * column entry)       (exact integer arithmetic, any order)
// is an upper bound of the row's fp64 score in units, and U[row] - slack a lower bound, slack = 2 per packed token
// + 5 per column token.
//   R  bm25_resolve_kernel  one record per (query, block, token): class and the term's run inside the block (range
//        table: no search) — a flat, fully parallel kernel pays the dependent look-ups once instead of every CTA.
//   A  bm25_filter_tma_kernel (block tiles of batched launches: posting runs staged through shared memory by a TMA
//        producer warp, rows owned by warps, 16-bit accumulators, 5 CTAs per SM — described at the kernel) or
//      bm25_filter_kernel<CH>  grid (tiles, queries), a tile = CH x 4096 rows (CH = 4: one block, used when the
//        launch has enough tiles to fill the GPU; CH = 1: a quarter block, for small corpora / single queries).
//        Run tokens: integer accumulators of the tile in shared memory; the CTA adds one token at a time, thread i
//        the i-th posting of the run (plain read-modify-write: the postings of one term are distinct rows), 8 loads
//        per thread in flight across rounds and tokens; an untabled (short) list is scanned whole.  Column tokens:
//        registers, 4096 rows at a time, merged with the shared accumulators.  Then the tile's H best allowed rows
//        by U ("heads") and the (H+1)-th best U ("rho": nothing that was not emitted is above it) go to global
//        memory: scores never leave the SM.  Selection by threshold: the (H+1)-th largest of a warp's 32 per-lane
//        maxima is reached by H+1 distinct rows, so only rows at or above it (H+1 and a few) can be among the
//        warp's H+1 best.
//   B  bm25_finish_kernel   one CTA per query: tau = k-th largest head minus the slack (k distinct rows reach it:
//        a valid lower bound of the k-th best exact score); every rho must be below it, else the query is flagged
//        (count = -1) and redone on the robust path; heads with U >= tau are the survivors (k + a handful): their
//        exact fp64 scores are recomputed from the postings — the same products, added in token order, so
//        bit-identical to numpy — and ordered by (score desc, row asc).
// Selective row filters (doc_filter keeping <= 4096 rows) skip the posting stream altogether: bm25_rows_kernel
// computes the exact scores of the listed rows only.
// ---------------------------------------------------------------------------
constexpr int kBmBlock = kBm25Block;  // rows per block: the granularity of the range tables and of the local row
constexpr int kBmQBits = 18;          // bits of q inside a packed posting
constexpr uint32_t kBmQMask = (1u << kBmQBits) - 1u;
constexpr int kBmMaxH = 31;           // heads per tile
constexpr int kBmList = 512;          // rows a CTA's warps hand to the final selection (8 warps x (H+1) and ties)
constexpr int kBmThetaHeads = 20;     // up to this many heads (+ rho) per tile the warps select by threshold
constexpr int kBmSurvivors = 1024;    // survivors per query the finish kernel re-scores
constexpr int kBmContrib = 4096;      // (survivor, token) products staged at a time
constexpr int kBmMaxQueryTokens = 1024;
constexpr int kBmDepth = 8;           // run postings per thread in flight (across rounds and tokens)
constexpr int kBmDnGroup = 2;         // column tokens whose loads are in flight together

// the tile's H+1 best of n_e (<= 32 * NU) list entries -> dst[0..H] (descending; 0 = none)
template <int NU>
__device__ __forceinline__ void bm25_emit_heads(const unsigned long long* s_list, int n_e, int H, int lane,
                                                unsigned long long* __restrict__ dst) {
    unsigned long long mine[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int e = lane + 32 * u;
        mine[u] = e < n_e ? s_list[e] : 0ull;
    }
    for (int h = 0; h <= H; ++h) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int u = 0; u < NU; ++u) m = mine[u] > m ? mine[u] : m;
        const uint32_t whi = __reduce_max_sync(0xffffffffu, (uint32_t)(m >> 32));
        const uint32_t wlo = __reduce_max_sync(0xffffffffu, (uint32_t)(m >> 32) == whi ? (uint32_t)m : 0u);
        const unsigned long long top = ((unsigned long long)whi << 32) | wlo;
        if (whi == 0u) {
            if (lane == 0)
                for (int hh = h; hh <= H; ++hh) dst[hh] = 0ull;
            break;
        }
#pragma unroll
        for (int u = 0; u < NU; ++u)
            if (mine[u] == top) mine[u] = 0ull;             // rows are distinct: exactly one entry
        if (lane == 0) dst[h] = top;
    }
}

// One record per (query, block, token slot), so that a tile's CTA starts with everything it needs in ONE load per
// thread instead of walking q_ptr -> terms -> term_info -> term_ptr -> rng_off itself: x = class, and
//   DENSE: y = column;   MID: [y, z) = the term's run inside this block;   LOW: [y, z) = the term's whole list.
// Slots past the query's last token are kBmSkip.
// low_search: an untabled (LOW) list is narrowed to its run inside the block by two searches and reported as MID, so
// that the consumer sees block-local runs only (bm25_filter_tma_kernel).
// the record of token slot i of (query q0 + q, block blk)
__device__ __forceinline__ uint4 bm25_resolve_one(const Bm25Device& ix, const int32_t* __restrict__ q_terms,
                                                  const int32_t* __restrict__ q_ptr, int q, int blk, int i, int low_search) {
    const int lo = q_ptr[q], nt = q_ptr[q + 1] - lo;
    uint4 d = make_uint4((uint32_t)kBmSkip, 0u, 0u, 0u);
    if (i < nt) {
        const int32_t t = q_terms[lo + i];
        if (t >= 0 && t < ix.n_terms) {
            const int2 info = ix.term_info[t];
            const int cls = (int)((unsigned)info.x >> 30);
            if (cls == kBmDense) {
                d = make_uint4((uint32_t)cls, (uint32_t)info.y, 0u, 0u);
            } else if (cls == kBmMid) {
                const uint32_t base = (uint32_t)ix.term_ptr[t];
                const int32_t* ro = ix.rng_off + (size_t)(info.x & 0x3FFFFFFF) * (ix.n_blocks + 1) + blk;
                d = make_uint4((uint32_t)cls, base + (uint32_t)ro[0], base + (uint32_t)ro[1], 0u);
            } else if (cls == kBmLow) {
                const int64_t lo_p = ix.term_ptr[t], hi_p = ix.term_ptr[t + 1];
                if (low_search) {
                    const int64_t a = bm25_lower_bound(ix.post_row, lo_p, hi_p, (int64_t)blk * kBmBlock);
                    const int64_t b = bm25_lower_bound(ix.post_row, a, hi_p, (int64_t)(blk + 1) * kBmBlock);
                    d = make_uint4((uint32_t)kBmMid, (uint32_t)a, (uint32_t)b, 0u);
                } else {
                    d = make_uint4((uint32_t)cls, (uint32_t)lo_p, (uint32_t)hi_p, 0u);
                }
            }
        }
    }
    return d;
}

__global__ void bm25_resolve_kernel(Bm25Device ix, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_ptr,
                                    int q0, int Q, int stride, int low_search, uint4* __restrict__ rec) {
    const int64_t per_q = (int64_t)ix.n_blocks * stride;
    const int64_t total = (int64_t)Q * per_q;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(e / per_q);
        const int rem = (int)(e - (int64_t)q * per_q);
        const int blk = rem / stride, i = rem - blk * stride;
        rec[e] = bm25_resolve_one(ix, q_terms, q_ptr, q0 + q, blk, i, low_search);
    }
}

// ---- column tokens, 4096 rows at a time: 4 x 8 bytes per thread and token straight into registers.  A thread
// owns the local rows c * 4096 + g * 1024 + 4 * tid + i (g, i = 0..3): 8-byte column loads and 16-byte
// shared accesses that are contiguous over the warp.  A column word holds two rows: s_all adds the words
// whole (sum of the low halves + 65536 * sum of the high halves, modulo 2^32), s_hi the high halves.  The
// sums are merged into the shared accumulators (the thread's own rows); the last pass also masks the
// disallowed rows and takes the thread's maximum.  Steps = (chunk, group of kBmDnGroup tokens); the loads
// of the next step are requested before the current one is added (two register buffers, ping-pong).
// xa holds step (chunk 0, group 0), requested by the caller (in flight while the run tokens are added).
template <int CH, int T, int DG = kBmDnGroup>
__device__ __forceinline__ void bm25_column_phase(uint32_t* acc, const uint16_t* const* s_colp, int n_col, bool last,
                                                  uint2 (&xa)[DG][4], const uint8_t* __restrict__ allow,
                                                  int64_t r0, int64_t r1, int tid, uint32_t& m) {
    constexpr int kTile = CH * 4096;
    constexpr int kChunk = 16 * T, kGrp = 4 * T, NCK = kTile / kChunk;
    const int ngroups = n_col > 0 ? (n_col + DG - 1) / DG : 1;
    uint32_t s_all[8], s_hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_all[j] = 0u; s_hi[j] = 0u; }
    int lc = 0, lg = 0;           // (chunk, group) of the next step to load
    int cc = 0, cg = 0;           // ... of the next step to add
    auto load_step = [&](uint2 (&xx)[DG][4]) {
        const bool live = lc < NCK;
#pragma unroll
        for (int u = 0; u < DG; ++u) {
            if (live && lg * DG + u < n_col) {
                const uint2* cp = reinterpret_cast<const uint2*>(s_colp[lg * DG + u] + lc * kChunk) + tid;
#pragma unroll
                for (int g = 0; g < 4; ++g) xx[u][g] = __ldg(cp + g * T);
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) xx[u][g] = make_uint2(0u, 0u);
            }
        }
        if (++lg == ngroups) { lg = 0; ++lc; }
    };
    auto add_step = [&](const uint2 (&xx)[DG][4]) {
#pragma unroll
        for (int u = 0; u < DG; ++u)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                s_all[2 * g + 0] += xx[u][g].x;
                s_hi[2 * g + 0] += xx[u][g].x >> 16;
                s_all[2 * g + 1] += xx[u][g].y;
                s_hi[2 * g + 1] += xx[u][g].y >> 16;
            }
        if (++cg == ngroups) {    // the chunk's last group: merge
            uint32_t* arow = acc + cc * kChunk + 4 * tid;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 a = *reinterpret_cast<const uint4*>(arow + g * kGrp);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t hi = s_hi[2 * g + h];
                    const uint32_t lo = s_all[2 * g + h] - (hi << 16);
                    (h ? a.z : a.x) += lo << kBmDenseShift;
                    (h ? a.w : a.y) += hi << kBmDenseShift;
                }
                if (last) {
                    if (allow != nullptr) {
                        const int64_t row = r0 + cc * kChunk + g * kGrp + 4 * tid;  // 4 rows inside one bitmap byte
                        const uint32_t bits = row < r1 ? ((uint32_t)allow[row >> 3] >> (row & 7)) : 0u;
                        if (!(bits & 1u)) a.x = 0u;
                        if (!(bits & 2u)) a.y = 0u;
                        if (!(bits & 4u)) a.z = 0u;
                        if (!(bits & 8u)) a.w = 0u;
                    }
                    const uint32_t m01 = a.x > a.y ? a.x : a.y, m23 = a.z > a.w ? a.z : a.w;
                    const uint32_t mg = m01 > m23 ? m01 : m23;
                    m = mg > m ? mg : m;
                }
                *reinterpret_cast<uint4*>(arow + g * kGrp) = a;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s_all[j] = 0u; s_hi[j] = 0u; }
            cg = 0;
            ++cc;
        }
    };
    uint2 xb[DG][4];
    if (n_col <= 3 * DG) {
        // the usual case, up to 6 column tokens: their pointers stay in registers and the group loop is
        // unrolled, so a step costs its loads and adds and little else
        const uint2* cp[3 * DG];
#pragma unroll
        for (int u = 0; u < 3 * DG; ++u)
            cp[u] = u < n_col ? reinterpret_cast<const uint2*>(s_colp[u]) + tid : nullptr;
        // requests step (chunk c, group G) into xx; G is a compile-time constant
#define BM25_LOAD_STEP(xx, G, c)                                                                        \
        do {                                                                                    \
            _Pragma("unroll") for (int u = 0; u < DG; ++u) {                            \
                if ((G) * DG + u < n_col) {                                             \
                    const uint2* q_ = cp[(G) * DG + u] + (c) * (kChunk / 4);                 \
                    _Pragma("unroll") for (int g = 0; g < 4; ++g) xx[u][g] = __ldg(q_ + g * T);   \
                } else {                                                                        \
                    _Pragma("unroll") for (int g = 0; g < 4; ++g) xx[u][g] = make_uint2(0u, 0u); \
                }                                                                               \
            }                                                                                   \
        } while (0)
#pragma unroll 1
        for (int c = 0; c < NCK; ++c) {
            // xa holds (c, 0)
            if (ngroups == 1) {
                if (c + 1 < NCK) BM25_LOAD_STEP(xb, 0, c + 1);
                add_step(xa);
            } else {
                BM25_LOAD_STEP(xb, 1, c);
                add_step(xa);
                if (ngroups == 2) {
                    if (c + 1 < NCK) BM25_LOAD_STEP(xa, 0, c + 1);
                    add_step(xb);
                    continue;
                }
                BM25_LOAD_STEP(xa, 2, c);
                add_step(xb);
                if (c + 1 < NCK) BM25_LOAD_STEP(xb, 0, c + 1);
                add_step(xa);
            }
            if (c + 1 < NCK) {
#pragma unroll
                for (int u = 0; u < DG; ++u)
#pragma unroll
                    for (int g = 0; g < 4; ++g) xa[u][g] = xb[u][g];
            }
        }
#undef BM25_LOAD_STEP
    } else {
        if (++lg == ngroups) { lg = 0; ++lc; }   // step 0 was requested before the run tokens
        while (cc < NCK) {
            load_step(xb);
            add_step(xa);
            if (cc >= NCK) break;
            load_step(xa);
            add_step(xb);
        }
    }
}

// ---- selection: the tile's H best allowed rows by U and the (H+1)-th best U -> dst[0..H] (descending; 0 = none).
// Every thread reads only the accumulators it wrote last (column phase): no barrier needed on entry.  *s_nlist is 0.
// NAMED: the CTA has a producer warp; the 256 consumer threads synchronise on named barrier 1.
// Column tokens of the 16-bit kernel (T = 256, one block).  The accumulators are 16 bits wide, two rows per 32-bit word
// (even row in the low half — the layout of a column word), in units coarser by 2^cshift than a packed posting's q: a
// column sum (entries = ceil(q / 4)) is divided by 2^(cshift - 2), rounding up, before it is added (still an upper
// bound; the caller widens the slack).  The same sums as bm25_column_phase, written for few instructions and registers — a column is addressed by a 32-bit offset (in 8-byte words) from the column base, so
// a step costs one shared load, one wide multiply-add, four 8-byte loads and their adds; the steps (chunk, token) are
// double buffered (xa holds step (0, 0), requested by the caller).
__device__ __forceinline__ void bm25_column_phase16(uint32_t* acc, const uint32_t* s_colo, const uint2* __restrict__ colbase,
                                                    int n_col, uint2 (&xa)[4], const uint8_t* __restrict__ allow,
                                                    int64_t r0, int64_t r1, int tid, int cshift, uint32_t& m) {
    constexpr int T = 256, kChunk = 16 * T, kGrp = 4 * T, NCK = kBmBlock / kChunk;
    const int cs = cshift - kBmDenseShift;
    const uint32_t rnd = (1u << cs) - 1u;
    const uint2* tb = colbase + tid;
    uint32_t s_all[8], s_hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_all[j] = 0u; s_hi[j] = 0u; }
    auto add = [&](const uint2 (&xx)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            s_all[2 * g + 0] += xx[g].x;
            s_hi[2 * g + 0] += xx[g].x >> 16;
            s_all[2 * g + 1] += xx[g].y;
            s_hi[2 * g + 1] += xx[g].y >> 16;
        }
    };
    auto load = [&](uint2 (&xx)[4], int u, int c) {
        const uint2* q_ = tb + (size_t)s_colo[u] + c * (kChunk / 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) xx[g] = __ldg(q_ + g * T);
    };
    auto merge = [&](int c) {
        uint32_t* arow = acc + ((c * kChunk + 4 * tid) >> 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint2 a = *reinterpret_cast<const uint2*>(arow + g * (kGrp / 2));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t hi = s_hi[2 * g + h];
                const uint32_t lo = s_all[2 * g + h] - (hi << 16);
                (h ? a.y : a.x) += ((lo + rnd) >> cs) | (((hi + rnd) >> cs) << 16);
                s_all[2 * g + h] = 0u;
                s_hi[2 * g + h] = 0u;
            }
            if (allow != nullptr) {
                const int64_t row = r0 + c * kChunk + g * kGrp + 4 * tid;      // 4 rows inside one bitmap byte
                const uint32_t bits = row < r1 ? ((uint32_t)allow[row >> 3] >> (row & 7)) : 0u;
                if (!(bits & 1u)) a.x &= 0xFFFF0000u;
                if (!(bits & 2u)) a.x &= 0x0000FFFFu;
                if (!(bits & 4u)) a.y &= 0xFFFF0000u;
                if (!(bits & 8u)) a.y &= 0x0000FFFFu;
            }
            const uint32_t x0 = a.x & 0xFFFFu, x1 = a.x >> 16, y0 = a.y & 0xFFFFu, y1 = a.y >> 16;
            const uint32_t m01 = x0 > x1 ? x0 : x1, m23 = y0 > y1 ? y0 : y1;
            const uint32_t mg = m01 > m23 ? m01 : m23;
            m = mg > m ? mg : m;
            *reinterpret_cast<uint2*>(arow + g * (kGrp / 2)) = a;
        }
    };
    if (n_col == 0) {
#pragma unroll 1
        for (int c = 0; c < NCK; ++c) merge(c);
        return;
    }
    uint2 xb[4];
    int u = 0, c = 0;                      // the step xa holds
#pragma unroll 1
    while (true) {
        // ---- xa is current, the next step goes to xb
        int nu = u + 1, nc = c;
        if (nu == n_col) { nu = 0; ++nc; }
        if (nc < NCK) load(xb, nu, nc);
        add(xa);
        if (nu == 0) merge(c);
        if (nc >= NCK) break;
        u = nu; c = nc;
        // ---- xb is current, the next step goes to xa
        nu = u + 1; nc = c;
        if (nu == n_col) { nu = 0; ++nc; }
        if (nc < NCK) load(xa, nu, nc);
        add(xb);
        if (nu == 0) merge(c);
        if (nc >= NCK) break;
        u = nu; c = nc;
    }
}

// the 4 accumulators at local rows [row, row + 4) (row a multiple of 4)
template <bool A16>
__device__ __forceinline__ void bm25_load4(const uint32_t* acc, uint32_t row, uint32_t (&v)[4]) {
    if constexpr (A16) {
        const uint2 a = *reinterpret_cast<const uint2*>(acc + (row >> 1));
        v[0] = a.x & 0xFFFFu; v[1] = a.x >> 16; v[2] = a.y & 0xFFFFu; v[3] = a.y >> 16;
    } else {
        const uint4 a = *reinterpret_cast<const uint4*>(acc + row);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
}

// A16: 16-bit accumulators in units of 2^ushift (see bm25_column_phase); the heads are written in the fine unit.
template <int CH, int T, bool NAMED, int LIST = kBmList, bool A16 = false>
__device__ __forceinline__ void bm25_select_phase(uint32_t* acc, uint32_t m, int H, int64_t r0, int tid,
                                                  unsigned long long* s_list, int* s_nlist_p,
                                                  unsigned long long* __restrict__ dst, int ushift = 0) {
    constexpr int kTile = CH * 4096;
    constexpr int kChunk = 16 * T, kGrp = 4 * T, NCK = kTile / kChunk;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t row_base = (uint32_t)r0 + 4u * (uint32_t)tid;
    if (H + 1 <= kBmThetaHeads) {
        // ---- theta = the (H+1)-th largest of the warp's 32 per-lane maxima: H+1 distinct rows reach it, so the
        // warp's H+1 best rows are among the rows >= theta (H+1 and a few)
        uint32_t theta = 0u;
        {
            uint32_t mc = m;
            for (int h = 0; h <= H; ++h) {
                theta = __reduce_max_sync(0xffffffffu, mc);
                if (theta == 0u) break;
                const unsigned who = __ballot_sync(0xffffffffu, mc == theta);
                if (lane == __ffs(who) - 1) mc = 0u;
            }
        }
        if (theta < 1u) theta = 1u;
        if (m >= theta) {
            // which of the thread's 4-row groups hold a row >= theta (a handful of rows per warp) ...
            uint32_t gmask = 0u;
#pragma unroll
            for (int j = 0; j < NCK * 4; ++j) {
                uint32_t a[4];
                bm25_load4<A16>(acc, (uint32_t)(j * kGrp + 4 * tid), a);
                const uint32_t m01 = a[0] > a[1] ? a[0] : a[1], m23 = a[2] > a[3] ? a[2] : a[3];
                gmask |= ((m01 > m23 ? m01 : m23) >= theta ? 1u : 0u) << j;
            }
            // ... and those rows
            while (gmask) {
                const int j = __ffs(gmask) - 1;
                gmask &= gmask - 1u;
                const uint32_t loc0 = (uint32_t)j * (uint32_t)kGrp;
                uint32_t vv[4];
                bm25_load4<A16>(acc, loc0 + 4u * (uint32_t)tid, vv);
                int slot = atomicAdd(s_nlist_p, (int)(vv[0] >= theta) + (int)(vv[1] >= theta) + (int)(vv[2] >= theta) + (int)(vv[3] >= theta));
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (vv[i] >= theta) {
                        if (slot < LIST)
                            s_list[slot] = ((unsigned long long)(vv[i] << ushift) << 32) | (unsigned long long)(~(row_base + loc0 + i));
                        ++slot;
                    }
            }
        }
    } else {
        // ---- many heads per tile (few tiles, or a large k): the warp's exact H+1 best, one per round (ties: lowest
        // row), straight from the shared accumulators
        for (int h = 0; h <= H; ++h) {
            uint32_t best = 0u, bloc = 0xFFFFFFFFu;
#pragma unroll 1
            for (int c = 0; c < NCK; ++c)
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t base = (uint32_t)(c * kChunk + g * kGrp);
                    uint32_t a[4];
                    bm25_load4<A16>(acc, base + 4u * (uint32_t)tid, a);
#pragma unroll
                    for (uint32_t i = 0; i < 4; ++i)
                        if (a[i] > best) { best = a[i]; bloc = base + i; }     // ascending rows, strict >: lowest row wins
                }
            const uint32_t wm = __reduce_max_sync(0xffffffffu, best);
            if (wm == 0u) break;
            const uint32_t myrow = best == wm ? row_base + bloc : 0xFFFFFFFFu;
            const uint32_t wr = __reduce_min_sync(0xffffffffu, myrow);
            if (myrow == wr) {
                if constexpr (A16) reinterpret_cast<uint16_t*>(acc)[bloc + 4 * tid] = (uint16_t)0;
                else acc[bloc + 4 * tid] = 0u;
                const int slot = atomicAdd(s_nlist_p, 1);
                if (slot < LIST) s_list[slot] = ((unsigned long long)(wm << ushift) << 32) | (unsigned long long)(~wr);
            }
        }
    }
    if (NAMED) asm volatile("bar.sync 1, 256;" ::: "memory"); else __syncthreads();
    if (warp == 0) {
        const int n_e = *s_nlist_p;
        if (n_e > LIST) {                                 // mass ties inside the tile: rho = "anything": redo
            for (int hh = lane; hh <= H; hh += 32) dst[hh] = ~0ull;
        } else if (n_e <= 64) {
            bm25_emit_heads<2>(s_list, n_e, H, lane, dst);
        } else if (n_e <= 256) {
            bm25_emit_heads<8>(s_list, n_e, H, lane, dst);
        } else {
            bm25_emit_heads<16>(s_list, n_e, H, lane, dst);
        }
    }
}

// CH: 4096-row units per tile; T: threads.  A thread owns 16 rows of every CHUNK of 16 * T rows (4 groups of 4 rows,
// 4 * T rows apart), so a tile is NCK = CH * 4096 / (16 * T) chunks.
template <int CH, int T>
__global__ void __launch_bounds__(T, T >= 512 ? 2 : 3)
bm25_filter_kernel(Bm25Device ix, const uint4* __restrict__ rec, int stride, const uint8_t* __restrict__ allow, int H,
                   unsigned long long* __restrict__ heads, const int32_t* __restrict__ q_terms,
                   const int32_t* __restrict__ q_ptr, int q0) {
    extern __shared__ __align__(16) uint32_t acc[];   // CH x 4096 accumulators
    __shared__ uint2 s_runs[kBmMaxTokens];            // [lo, hi) inside the packed stream: tabled runs from the front,
    __shared__ const uint16_t* s_colp[kBmMaxTokens];  // scanned lists from the back; column of the tile per DENSE token
    __shared__ int s_ntab, s_nscan, s_ncol, s_nlist;
    __shared__ unsigned long long s_list[kBmList];
    constexpr int kTile = CH * 4096;
    constexpr int kChunk = 16 * T, kGrp = 4 * T, NCK = kTile / kChunk;
    static_assert(NCK >= 1 && NCK * kChunk == kTile, "tile must be whole chunks");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    constexpr int kSubBits = CH == 4 ? 0 : (CH == 2 ? 1 : 2);      // tiles per block = 1 << kSubBits
    const int blk = tile >> kSubBits;
    const uint32_t quarter = (uint32_t)(tile & ((1 << kSubBits) - 1));   // CH < 4: which part of the block
    const int64_t r0 = (int64_t)tile * kTile;
    const int64_t r1 = r0 + kTile < ix.n_docs ? r0 + kTile : ix.n_docs;
    const uint4* my_rec = rec + ((size_t)blockIdx.y * ix.n_blocks + blk) * stride;
#pragma unroll
    for (int j = 0; j < NCK * 4; ++j) *reinterpret_cast<uint4*>(&acc[j * kGrp + 4 * tid]) = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) s_nlist = 0;
    uint32_t m = 0u;                      // the thread's largest upper bound (set in the last pass)
    for (int t0 = 0; t0 < stride; t0 += kBmMaxTokens) {
        const int tn = stride - t0 < kBmMaxTokens ? stride - t0 : kBmMaxTokens;
        const bool last = t0 + kBmMaxTokens >= stride;
        uint4 d = make_uint4((uint32_t)kBmSkip, 0u, 0u, 0u);
        // (rec == nullptr: small launches — a single query — resolve their records here and save the resolve launch)
        if (tid < tn) d = rec != nullptr ? __ldg(my_rec + t0 + tid)
                                         : bm25_resolve_one(ix, q_terms, q_ptr, q0 + (int)blockIdx.y, blk, t0 + tid, 0);
        __syncthreads();                  // accumulators zeroed / previous pass done with the token table
        if (tid == 0) { s_ntab = 0; s_nscan = 0; s_ncol = 0; }
        __syncthreads();
        if (d.x == (uint32_t)kBmDense) {
            s_colp[atomicAdd(&s_ncol, 1)] =
                ix.dense_col + ((size_t)d.y * ix.n_blocks + blk) * kBmBlock + quarter * (uint32_t)kTile;
        } else if (d.x == (uint32_t)kBmMid) {
            if (d.z > d.y) s_runs[atomicAdd(&s_ntab, 1)] = make_uint2(d.y, d.z);
        } else if (d.x == (uint32_t)kBmLow) {
            if (d.z > d.y) s_runs[kBmMaxTokens - 1 - atomicAdd(&s_nscan, 1)] = make_uint2(d.y, d.z);
        }
        __syncthreads();
        const int n_tab = s_ntab, n_run = n_tab + s_nscan, n_col = s_ncol;
        // ---- run tokens, one at a time for the whole CTA: thread i adds the i-th, 256+i-th, ... posting of the run
        // (the postings of one term are distinct rows: plain read-modify-write), a barrier separates the terms.
        // The (token, round) sequence is the same for every thread; kBmDepth of its loads are in flight.
        // An untabled list is scanned whole: only the rows of this tile count.
        // the first column step (chunk 0, first group) is requested now: in flight while the run tokens are added
        uint2 xa[kBmDnGroup][4];
#pragma unroll
        for (int u = 0; u < kBmDnGroup; ++u) {
            if (u < n_col) {
                const uint2* cp = reinterpret_cast<const uint2*>(s_colp[u]) + tid;
#pragma unroll
                for (int g = 0; g < 4; ++g) xa[u][g] = __ldg(cp + g * T);
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) xa[u][g] = make_uint2(0u, 0u);
            }
        }
        {
            // iterator over (token, round), block-uniform: the runs in s_runs are non-empty
            int it_e = 0;
            bool it_scan = n_tab == 0;
            uint2 it_r = n_run > 0 ? s_runs[it_scan ? kBmMaxTokens - 1 : 0] : make_uint2(0u, 0u);
            uint32_t it_p = it_r.x;
            bool it_first = true;
            uint32_t val[kBmDepth];
            bool first[kBmDepth], ok[kBmDepth];
            // fills slot dd with the thread's posting of the current item and moves on
#define BM25_FETCH(dd)                                                                              \
            do {                                                                                    \
                ok[dd] = it_e < n_run;                                                              \
                first[dd] = it_first;                                                               \
                val[dd] = 0u;                     /* 0: adds nothing (a posting's q is >= 2) */     \
                if (ok[dd]) {                                                                       \
                    const uint32_t p = it_p + (uint32_t)tid;                                        \
                    if (p < it_r.y) {                                                               \
                        val[dd] = __ldg(ix.post_pack + p);                                          \
                        if (it_scan) {                                                              \
                            const int64_t row = ix.post_row[p];                                     \
                            if (row < r0 || row >= r1) val[dd] = 0u;                                \
                        }                                                                           \
                    }                                                                               \
                    it_p += T;                                                                      \
                    it_first = false;                                                               \
                    if (it_p >= it_r.y) {         /* next token */                                  \
                        ++it_e;                                                                     \
                        it_first = true;                                                            \
                        if (it_e < n_run) {                                                         \
                            it_scan = it_e >= n_tab;                                                \
                            it_r = s_runs[it_scan ? kBmMaxTokens - 1 - (it_e - n_tab) : it_e];      \
                            it_p = it_r.x;                                                          \
                        }                                                                           \
                    }                                                                               \
                }                                                                                   \
            } while (0)
#pragma unroll
            for (int dd = 0; dd < kBmDepth; ++dd) BM25_FETCH(dd);
            bool any = false;
            while (ok[0]) {               // block-uniform; slots are consumed in the order they were filled
#pragma unroll
                for (int dd = 0; dd < kBmDepth; ++dd) {
                    if (ok[dd]) {
                        if (first[dd] && any) __syncthreads();       // the previous term is done
                        any = true;
                        if (val[dd]) {
                            const uint32_t loc = val[dd] >> kBmQBits;                  // row inside the block
                            if (CH == 4) acc[loc] += val[dd] & kBmQMask;
                            else if ((loc / (uint32_t)kTile) == quarter) acc[loc % (uint32_t)kTile] += val[dd] & kBmQMask;
                        }
                        BM25_FETCH(dd);
                    }
                }
            }
#undef BM25_FETCH
        }
        __syncthreads();
        if (n_col > 0 || last) bm25_column_phase<CH, T>(acc, s_colp, n_col, last, xa, allow, r0, r1, tid, m);
    }
    bm25_select_phase<CH, T, false>(acc, m, H, r0, tid, s_list, &s_nlist,
                                    heads + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (H + 1));
}

// ---------------------------------------------------------------------------
// bm25_filter_tma_kernel — the block-tile filter (16384 rows per CTA) with the posting runs STAGED THROUGH SHARED
// MEMORY by a producer warp: every run of the tile is cut into chunks of <= 512 postings, copied with TMA 1-D bulk
// copies (cp.async.bulk + mbarrier complete_tx) into a 4-slot ring, so that the loads hold no registers, need no
// per-thread address arithmetic and run ahead of the consumers across terms.  The 8 consumer warps OWN 2048 rows of
// the tile each: a chunk is sorted by row (the packed word carries the local row in its top 14 bits), so a warp
// finds its sub-run with a two-level warp-parallel search in shared memory (2 loads, 2 ballots) and adds it with
// plain read-modify-writes — different terms can only meet inside one warp, which takes them in program order, so
// there is no CTA-wide barrier between terms.  Column tokens and the selection are bm25_filter_kernel's.
// CTAs are launched in (block, query) order: the CTAs resident together share a block's columns and runs in L2.
// Measured at config 4 (0.58 ms per 256 queries): 2 CTAs per SM with 10- or 20-slot rings 0.73 ms, 2 CTAs of 16 consumer
// warps 0.82 ms, a 5th slot or an L2 prefetch of the runs no change (resident warps matter, ring depth does not); a
// persistent variant whose producer runs one tile ahead (second token table, work counter) 0.60 ms: dropped.
// Needs stride <= kBmMaxTokens and block-local runs for every term (bm25_resolve_kernel with low_search).
// ---------------------------------------------------------------------------
constexpr int kBmRingSlots = 4;
constexpr int kBmSlotPost = 512;                  // postings per chunk
constexpr int kBmSlotWords = kBmSlotPost + 4;     // the copy starts / ends on 16-byte boundaries of the packed stream
constexpr int kBmTmaThreads = 288;                // 8 consumer warps + the producer warp
constexpr size_t bm25_tma_smem(bool a16) {
    return (size_t)kBmBlock * (a16 ? 2 : 4) + (size_t)kBmRingSlots * kBmSlotWords * 4;
}
static_assert((size_t)kBmList * 8 <= (size_t)kBmRingSlots * kBmSlotWords * 4, "the selection list reuses the ring");

// A16: 16-bit accumulators (two rows per word) in units of 2^cshift: 32 KB instead of 64 KB per tile, so FOUR CTAs are
// resident per SM instead of three (the kernel is bound by resident warps, not by bytes).  A posting adds
// ceil(q / 2^cshift); the caller picks cshift so that the sum of a query's tokens cannot reach 2^16 and widens the
// finish kernel's slack by 2^cshift per token.
template <bool A16>
__global__ void __launch_bounds__(kBmTmaThreads, A16 ? 5 : 3)
bm25_filter_tma_kernel(Bm25Device ix, const uint4* __restrict__ rec, int stride, const uint8_t* __restrict__ allow, int H,
                       int by_block, int cshift, unsigned long long* __restrict__ heads) {
    extern __shared__ __align__(16) uint32_t acc[];   // 16384 accumulators | ring
    uint32_t* ring = acc + (A16 ? kBmBlock / 2 : kBmBlock);
    __shared__ uint2 s_runs[kBmMaxTokens];            // [lo, hi) inside the packed stream
    __shared__ const uint16_t* s_colp[kBmMaxTokens];  // column of the tile per DENSE token
    __shared__ uint32_t s_colo[kBmMaxTokens];         // A16: the same as an offset from the column base, in 8-byte words
    __shared__ __align__(8) uint64_t s_full[kBmRingSlots], s_empty[kBmRingSlots];
    __shared__ int s_nrun, s_ncol, s_nlist;
    constexpr int T = 256;
    constexpr int DG = A16 ? 1 : kBmDnGroup;          // column tokens in flight together (A16: 40 registers, 5 CTAs/SM)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // grid (queries, blocks) or (blocks, queries): CTAs are dealt x-fastest, so `by_block` runs all queries of one
    // block back to back — the block's columns and runs are then shared in L2 by the CTAs that are resident together
    const int blk = by_block ? blockIdx.y : blockIdx.x;
    const int qy = by_block ? blockIdx.x : blockIdx.y;
    const int64_t r0 = (int64_t)blk * kBmBlock;
    const int64_t r1 = r0 + kBmBlock < ix.n_docs ? r0 + kBmBlock : ix.n_docs;
    const uint4* my_rec = rec + ((size_t)qy * ix.n_blocks + blk) * stride;
    uint4 d = make_uint4((uint32_t)kBmSkip, 0u, 0u, 0u);
    if (tid < stride) d = __ldg(my_rec + tid);
    if (tid < T) {
#pragma unroll
        for (int j = 0; j < (A16 ? 8 : 16); ++j)
            *reinterpret_cast<uint4*>(&acc[j * 4 * T + 4 * tid]) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        s_nrun = 0; s_ncol = 0; s_nlist = 0;
        for (int s = 0; s < kBmRingSlots; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 8); }
        mbar_fence_init();
    }
    __syncthreads();
    if (d.x == (uint32_t)kBmDense) {
        const int slot_c = atomicAdd(&s_ncol, 1);
        s_colp[slot_c] = ix.dense_col + ((size_t)d.y * ix.n_blocks + blk) * kBmBlock;
        s_colo[slot_c] = (uint32_t)(((size_t)d.y * ix.n_blocks + blk) * (kBmBlock / 4));
    } else if (d.x == (uint32_t)kBmMid) {
        if (d.z > d.y) s_runs[atomicAdd(&s_nrun, 1)] = make_uint2(d.y, d.z);
    }
    __syncthreads();
    const int n_run = s_nrun, n_col = s_ncol;
    const uint32_t full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]), ring0 = smem_u32(ring);
    if (warp == 8) {
        // ===================== producer: every chunk of every run, in list order =====================
        if (lane == 0) {
            // the tile's columns are asked into L2 first: the consumers' register loads then find them there
            for (int u = 0; u < n_col; ++u)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(s_colp[u]), "r"(kBmBlock * 2) : "memory");
            int slot = 0;
            uint32_t ph = 0;
            for (int r = 0; r < n_run; ++r) {
                const uint2 run = s_runs[r];
                for (uint32_t p = run.x; p < run.y; p += kBmSlotPost) {
                    const uint32_t e = p + kBmSlotPost < run.y ? p + kBmSlotPost : run.y;
                    const uint32_t a0 = p & ~3u, a1 = (e + 3u) & ~3u;
                    mbar_wait_a(empty0 + 8u * slot, ph ^ 1u);
                    mbar_arrive_expect_tx_a(full0 + 8u * slot, (a1 - a0) * 4u);
                    bulk_g2s_a(ring0 + slot * (kBmSlotWords * 4), ix.post_pack + a0, (a1 - a0) * 4u, full0 + 8u * slot);
                    if (++slot == kBmRingSlots) { slot = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }
    // ===================== consumers =====================
    // the first column step (chunk 0, first group) is requested now: in flight while the run tokens are added
    uint2 xa[DG][4];
#pragma unroll
    for (int u = 0; u < DG; ++u) {
        if (u < n_col) {
            const uint2* cp = reinterpret_cast<const uint2*>(s_colp[u]) + tid;
#pragma unroll
            for (int g = 0; g < 4; ++g) xa[u][g] = __ldg(cp + g * T);
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) xa[u][g] = make_uint2(0u, 0u);
        }
    }
    {
        const uint32_t t_lo = (uint32_t)warp << 29;           // first local row of the warp, in the packed word's place
        const uint32_t t_hi = (uint32_t)(warp + 1) << 29;     // (warp 7: wraps to 0, its upper bound is the chunk's end)
        uint16_t* acc16 = reinterpret_cast<uint16_t*>(acc);
        const uint32_t rnd = (1u << cshift) - 1u;
        // adds one posting to its row's accumulator
        auto add = [&](uint32_t v) {
            if constexpr (A16) {
                const uint32_t loc = v >> kBmQBits;
                acc16[loc] = (uint16_t)(acc16[loc] + (((v & kBmQMask) + rnd) >> cshift));
            } else {
                acc[v >> kBmQBits] += v & kBmQMask;
            }
        };
        int slot = 0;
        uint32_t ph = 0;
        for (int r = 0; r < n_run; ++r) {
            const uint2 run = s_runs[r];
            for (uint32_t p = run.x; p < run.y; p += kBmSlotPost) {
                const int n = (int)((p + kBmSlotPost < run.y ? p + kBmSlotPost : run.y) - p);
                const uint32_t* c = ring + slot * kBmSlotWords + (p & 3u);
                mbar_wait_a(full0 + 8u * slot, ph);
                if (n <= 64) {
                    // short chunk: every lane looks at (up to) two postings and adds the ones its warp owns
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = lane + 32 * u;
                        if (i < n) {
                            const uint32_t v = c[i];
                            if ((v >> 29) == (uint32_t)warp) add(v);
                        }
                    }
                } else {
                    // level 1: probes 16 apart; level 2: the 16 postings behind the last probe below the target
                    // (lanes 0-15 for the lower bound, 16-31 for the upper one)
                    const int i1 = lane * 16;
                    const uint32_t v1 = i1 < n ? c[i1] : 0xFFFFFFFFu;
                    const int s_lo = __popc(__ballot_sync(0xffffffffu, v1 < t_lo));
                    const int s_hi = warp == 7 ? 0 : __popc(__ballot_sync(0xffffffffu, v1 < t_hi));
                    const bool up = lane >= 16;
                    const int sb = up ? s_hi : s_lo;
                    const int i2 = (sb - 1) * 16 + 1 + (lane & 15);
                    const uint32_t v2 = (sb > 0 && i2 < n) ? c[i2] : 0xFFFFFFFFu;
                    const unsigned b2 = __ballot_sync(0xffffffffu, v2 < (up ? t_hi : t_lo));
                    const int lo = s_lo > 0 ? (s_lo - 1) * 16 + 1 + __popc(b2 & 0xFFFFu) : 0;
                    const int hi = warp == 7 ? n : (s_hi > 0 ? (s_hi - 1) * 16 + 1 + __popc(b2 >> 16) : 0);
                    for (int i = lo + lane; i < hi; i += 32) add(c[i]);
                }
                __syncwarp();                                  // the next chunk may be another term on the same rows
                if (lane == 0) mbar_arrive_a(empty0 + 8u * slot);
                if (++slot == kBmRingSlots) { slot = 0; ph ^= 1u; }
            }
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");             // the column phase owns rows by thread
    uint32_t m = 0u;
    if constexpr (A16)
        bm25_column_phase16(acc, s_colo, reinterpret_cast<const uint2*>(ix.dense_col), n_col, xa[0], allow, r0, r1, tid,
                            cshift, m);
    else
        bm25_column_phase<4, T, DG>(acc, s_colp, n_col, true, xa, allow, r0, r1, tid, m);
    bm25_select_phase<4, T, true, kBmList, A16>(acc, m, H, r0, tid, reinterpret_cast<unsigned long long*>(ring), &s_nlist,
                                                heads + ((size_t)qy * ix.n_blocks + blk) * (H + 1), A16 ? cshift : 0);
}

// w * impact of (term t, row) if the posting exists, else 0: the product the accumulation adds
__device__ __forceinline__ double bm25_term_contrib(const Bm25Device& ix, int32_t t, uint32_t row) {
    if (t < 0 || t >= ix.n_terms) return 0.0;
    const double w = ix.idf[t];
    if (w == 0.0) return 0.0;
    int64_t lo = ix.term_ptr[t], hi = ix.term_ptr[t + 1];
    if (ix.term_info) {
        const int2 info = ix.term_info[t];
        const int cls = (int)((unsigned)info.x >> 30);
        if (cls == kBmMid || cls == kBmDense) {            // the range table narrows the search to the row's range
            const int32_t* ro = ix.rng_off + (size_t)(info.x & 0x3FFFFFFF) * (ix.n_blocks + 1) + (row / (uint32_t)kBmBlock);
            hi = lo + ro[1];
            lo = lo + ro[0];
        }
    }
    const int64_t pos = bm25_lower_bound(ix.post_row, lo, hi, (int64_t)row);
    if (pos < hi && (uint32_t)ix.post_row[pos] == row) return __dmul_rn(w, ix.post_impact[pos]);
    return 0.0;
}

// Block-wide: exact fp64 scores of n rows (token order, numpy's `score +=`), then the k best with score > 0 by
// (score desc, row asc).  rows: shared or global memory; contrib: kBmContrib doubles; keys: the next power of
// two >= max(n, 32) entries.  All threads of the 256-thread CTA call it.
__device__ __forceinline__ void bm25_exact_topk(const Bm25Device& ix, const int32_t* __restrict__ terms, int nt,
                                                const uint32_t* rows, int n, int k, double* contrib, Bm25Key* keys,
                                                int32_t* out_rows, double* out_scores, int32_t* out_count) {
    const int cs = nt > 0 ? (kBmContrib / nt > 0 ? kBmContrib / nt : 1) : (n > 0 ? n : 1);
    for (int s0 = 0; s0 < n; s0 += cs) {
        const int csn = n - s0 < cs ? n - s0 : cs;
        __syncthreads();
        for (int pair = threadIdx.x; pair < csn * nt; pair += blockDim.x) {
            const int s = pair / nt, i = pair - s * nt;
            contrib[pair] = bm25_term_contrib(ix, terms[i], rows[s0 + s]);
        }
        __syncthreads();
        for (int s = threadIdx.x; s < csn; s += blockDim.x) {
            double sc = 0.0;
            for (int i = 0; i < nt; ++i) sc = __dadd_rn(sc, contrib[s * nt + i]);     // + 0.0 where no posting
            Bm25Key key{0ull, 0u, 0u};
            if (sc > 0.0) { key.s = (uint64_t)__double_as_longlong(sc); key.nrow = ~rows[s0 + s]; }
            keys[s0 + s] = key;
        }
    }
    __syncthreads();
    if (n <= (int)blockDim.x) {
        // the usual case (k + a handful of survivors): order by rank counting, no sort
        Bm25Key me{0ull, 0u, 0u};
        if ((int)threadIdx.x < n) me = keys[threadIdx.x];
        const int valid = __syncthreads_count(me.s != 0ull);
        if (me.s != 0ull) {                                   // (the warps past the n-th survivor have nothing to rank)
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += me < keys[j];
            if (rank < k) {
                out_rows[rank] = (int32_t)(~me.nrow);
                out_scores[rank] = __longlong_as_double((long long)me.s);
            }
        }
        const int nout = valid < k ? valid : k;
        for (int i = nout + threadIdx.x; i < k; i += blockDim.x) { out_rows[i] = -1; out_scores[i] = 0.0; }
        if (threadIdx.x == 0) *out_count = nout;
        return;
    }
    int nsort = 32;
    while (nsort < n) nsort <<= 1;
    for (int i = n + threadIdx.x; i < nsort; i += blockDim.x) keys[i] = Bm25Key{0ull, 0u, 0u};
    block_bitonic_desc(keys, nsort);
    int local = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < nsort && keys[i].s != 0ull;
        local += ok;
        out_rows[i] = ok ? (int32_t)(~keys[i].nrow) : -1;
        out_scores[i] = ok ? __longlong_as_double((long long)keys[i].s) : 0.0;
    }
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) *out_count = s_cnt;
}

constexpr int kBmFinishThreads = 1024;   // (survivor, token) look-ups are chains of dependent loads: all of them at once
__global__ void __launch_bounds__(kBmFinishThreads)
bm25_finish_kernel(Bm25Device ix, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_ptr, int q0,
                   const unsigned long long* __restrict__ heads, int n_tiles, int H, int h_tau, int coarse_shift, int k,
                   int32_t* out_rows, double* out_scores, int32_t* out_counts) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    // [ 32 KB: tau sort buffer, later the staged products | survivors' keys | survivors' rows ]
    double* contrib = reinterpret_cast<double*>(sm_raw);
    Bm25Key* keys = reinterpret_cast<Bm25Key*>(sm_raw + (size_t)kBmContrib * sizeof(double));
    uint32_t* surv = reinterpret_cast<uint32_t*>(keys + kBmSurvivors);
    __shared__ int s_n, s_flag, s_extra;
    __shared__ uint32_t s_hist[256], s_sel[2];
    const int q = blockIdx.x;
    const int n_ranges = n_tiles;
    const unsigned long long* hq = heads + (size_t)q * n_ranges * (H + 1);
    const int32_t* terms = q_terms + q_ptr[q0 + q];
    const int nt = q_ptr[q0 + q + 1] - q_ptr[q0 + q];
    if (threadIdx.x == 0) { s_n = 0; s_flag = 0; s_extra = 0; }
    // tau: the k-th largest of the first h_tau heads of every range (distinct rows), minus the slack
    const int n_tau = n_ranges * h_tau;
    // (radix select, one byte per pass from the top: 12 barriers instead of the ~55 of a block sort of 1024 heads)
    uint32_t* su = reinterpret_cast<uint32_t*>(sm_raw);       // the n_tau upper bounds (<= 4096)
    for (int e = threadIdx.x; e < n_tau; e += blockDim.x)
        su[e] = (uint32_t)(hq[(size_t)(e / h_tau) * (H + 1) + (e % h_tau)] >> 32);
    uint32_t tau_u = 0u;
    if (k <= n_tau) {
        uint32_t prefix = 0u, kr = (uint32_t)k;               // the kr-th largest of the entries that match the prefix
        for (int sh = 24; sh >= 0; sh -= 8) {
            if (threadIdx.x < 256) s_hist[threadIdx.x] = 0u;
            __syncthreads();
            const uint32_t himask = sh == 24 ? 0u : (0xFFFFFFFFu << (sh + 8));
            for (int e = threadIdx.x; e < n_tau; e += blockDim.x) {
                const uint32_t u = su[e];
                if ((u & himask) == prefix) atomicAdd(&s_hist[(u >> sh) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {                           // lane l owns the bins [8l, 8l + 8)
                const int lane = threadIdx.x;
                uint32_t c[8], sum = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = s_hist[8 * lane + j]; sum += c[j]; }
                uint32_t incl = sum;                          // entries in this lane's bins and above
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
                    if (lane + o < 32) incl += t;
                }
                uint32_t a = incl - sum;                      // entries above this lane's bins
                if (a < kr && kr <= a + sum) {
#pragma unroll
                    for (int j = 7; j >= 0; --j) {
                        if (a + c[j] >= kr) { s_sel[0] = prefix | ((uint32_t)(8 * lane + j) << sh); s_sel[1] = kr - a; break; }
                        a += c[j];
                    }
                }
            }
            __syncthreads();
            prefix = s_sel[0];
            kr = s_sel[1];
        }
        tau_u = prefix;
    }
    __syncthreads();
    // U - slack <= score / unit <= U: a packed posting is at most 2 units above its product, a column entry at
    // most 2 + 15 (tokens beyond kBmMaxQueryTokens: the query is redone on the robust path anyway)
    for (int i = threadIdx.x; i < nt && i < kBmMaxQueryTokens; i += blockDim.x) {
        const int32_t t = terms[i];
        if (t >= 0 && t < ix.n_terms && (int)((unsigned)ix.term_info[t].x >> 30) == kBmDense) atomicAdd(&s_extra, 1);
    }
    __syncthreads();
    // (16-bit accumulators: every token was rounded up to the coarse unit once more)
    const uint32_t slack = 2u * (uint32_t)nt + (uint32_t)((1 << kBmDenseShift) - 1) * (uint32_t)s_extra + 2u +
                           (coarse_shift > 0 ? ((uint32_t)nt << coarse_shift) : 0u);
    uint32_t thr = tau_u > slack ? tau_u - slack : 0u;
    if (thr < 1u) thr = 1u;                                  // no bound: every row with a positive upper bound
    __syncthreads();
    for (int e = threadIdx.x; e < n_ranges * (H + 1); e += blockDim.x) {
        const unsigned long long key = hq[e];
        const uint32_t u = (uint32_t)(key >> 32);
        if (u < thr) continue;
        if (e % (H + 1) == H) {
            s_flag = 1;                                      // a row that was NOT emitted may reach tau
        } else {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < kBmSurvivors) surv[slot] = ~(uint32_t)key;
        }
    }
    __syncthreads();
    const int n = s_n;
    if (s_flag || n > kBmSurvivors || nt > kBmMaxQueryTokens) {      // redo on the robust path
        if (threadIdx.x == 0) out_counts[q] = -1;
        return;
    }
    bm25_exact_topk(ix, terms, nt, surv, n, k, contrib, keys, out_rows + (size_t)q * k, out_scores + (size_t)q * k,
                    out_counts + q);
}

// selective row filter: the exact scores of the listed rows, nothing else is read
__global__ void __launch_bounds__(256)
bm25_rows_kernel(Bm25Device ix, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_ptr,
                 const int32_t* __restrict__ rows, int n_rows, int k, int32_t* out_rows, double* out_scores,
                 int32_t* out_counts) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    double* contrib = reinterpret_cast<double*>(sm_raw);
    Bm25Key* keys = reinterpret_cast<Bm25Key*>(sm_raw + (size_t)kBmContrib * sizeof(double));
    const int q = blockIdx.x;
    const int32_t* terms = q_terms + q_ptr[q];
    int nt = q_ptr[q + 1] - q_ptr[q];
    // more tokens than the staging area holds per row: take them in the kernel's stride (cs = 1 handles any nt
    // up to kBmContrib; beyond that the host splits the call)
    bm25_exact_topk(ix, terms, nt, reinterpret_cast<const uint32_t*>(rows), n_rows, k, contrib, keys,
                    out_rows + (size_t)q * k, out_scores + (size_t)q * k, out_counts + q);
}

// ---------------------------------------------------------------------------
// index-time kernels of the fast path
// ---------------------------------------------------------------------------
__global__ void bm25_maximpact_kernel(const double* __restrict__ impact, int64_t nnz, unsigned long long* out) {
    double m = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const double v = impact[p];
        m = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

cudaError_t bm25_cmax_launch(const Bm25Device& ix, unsigned long long* cmax, cudaStream_t st) {
    int64_t g = (ix.nnz + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    bm25_maximpact_kernel<<<(int)g, 256, 0, st>>>(ix.post_impact, ix.nnz, cmax);
    return cudaGetLastError();
}

__global__ void bm25_pack_kernel(Bm25Device ix, double inv_unit) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ix.nnz; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = ix.n_terms;                     // the term of posting p: last t with term_ptr[t] <= p
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (ix.term_ptr[mid] <= p) lo = mid; else hi = mid;
        }
        const double c = __dmul_rn(ix.idf[lo], ix.post_impact[p]);
        uint32_t q = 0u;
        if (c > 0.0) {
            const double v = ceil(c * inv_unit) + 1.0;       // q * unit >= c and (q - 2) * unit <= c
            q = v > (double)kBmQMask ? kBmQMask : (uint32_t)v;
        }
        ix.post_pack[p] = (((uint32_t)ix.post_row[p] & (uint32_t)(kBmBlock - 1)) << kBmQBits) | q;
    }
}

__global__ void bm25_rng_kernel(Bm25Device ix, const int32_t* __restrict__ tabled_terms, int n_tabled) {
    const int64_t total = (int64_t)n_tabled * (ix.n_blocks + 1);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int slot = (int)(e / (ix.n_blocks + 1)), r = (int)(e % (ix.n_blocks + 1));
        const int32_t t = tabled_terms[slot];
        const int64_t lo = ix.term_ptr[t], hi = ix.term_ptr[t + 1];
        ix.rng_off[e] = (int32_t)(bm25_lower_bound(ix.post_row, lo, hi, (int64_t)r * kBmBlock) - lo);
    }
}

// column of a DENSE term: entry (row) = ceil(q / 16) of the term's posting on that row (columns zeroed by the caller)
__global__ void bm25_column_kernel(Bm25Device ix, const int32_t* __restrict__ dense_terms) {
    const int32_t t = dense_terms[blockIdx.y];
    const int64_t lo = ix.term_ptr[t], hi = ix.term_ptr[t + 1];
    uint16_t* col = ix.dense_col + (size_t)blockIdx.y * ((size_t)ix.n_blocks * kBmBlock);
    for (int64_t p = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hi; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t q = ix.post_pack[p] & kBmQMask;
        col[ix.post_row[p]] = (uint16_t)((q + ((1u << kBmDenseShift) - 1u)) >> kBmDenseShift);
    }
}

cudaError_t bm25_index_build_launch(const Bm25Device& ix, double unit, const int32_t* tabled_terms, int n_tabled,
                                    const int32_t* dense_terms, int n_dense, cudaStream_t st) {
    auto grid_for = [](int64_t n) {
        int64_t g = (n + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        return (int)(g < 1 ? 1 : g);
    };
    if (ix.nnz > 0) {
        bm25_pack_kernel<<<grid_for(ix.nnz), 256, 0, st>>>(ix, 1.0 / unit);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (n_tabled > 0) {
        bm25_rng_kernel<<<grid_for((int64_t)n_tabled * (ix.n_blocks + 1)), 256, 0, st>>>(ix, tabled_terms, n_tabled);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (n_dense > 0) {
        dim3 grid(148, n_dense);
        bm25_column_kernel<<<grid, 256, 0, st>>>(ix, dense_terms);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// heads per tile: enough that a tile holding more than H of the query's top rows is a < 1e-3 event when the top rows
// fall into tiles independently (lambda = expected top rows per tile)
static int bm25_heads_per_tile(int n_tiles, int k) {
    const double lambda = 1.15 * k / n_tiles;
    int H = (int)ceil(lambda + 4.0 * sqrt(lambda) + 4.0);
    return H < 4 ? 4 : H;
}
// heads per tile that enter the threshold: the k-th largest of n_tiles x h_tau heads (distinct rows) is a lower bound
// of the k-th best score for any h_tau >= k / n_tiles.  Many tiles: the k-th largest tile maximum is tight enough.
// Few tiles: with k / n_tiles heads the k-th largest sits at the weakest tile's last head (a ragged last tile
// drags it far down and every tile then holds more than H rows above it), so take about 2k heads in total.
static int bm25_tau_heads(int n_tiles, int k) {
    const int need = (k + n_tiles - 1) / n_tiles;
    if (n_tiles >= 4 * k) return need;
    const int want = (2 * k + n_tiles - 1) / n_tiles + 2;
    const int H = bm25_heads_per_tile(n_tiles, k);
    return want < H ? (want > need ? want : need) : (H > need ? H : need);
}

struct Bm25Plan {
    int ch;          // 4096-row chunks per tile: 4 (a whole block) or 1
    int n_tiles, H, h_tau;
};
static bool bm25_plan_for(const Bm25Device& ix, int k, int ch, Bm25Plan* p) {
    p->ch = ch;
    p->n_tiles = (int)((ix.n_docs + (int64_t)ch * 4096 - 1) / ((int64_t)ch * 4096));
    if (p->n_tiles < 1) return false;
    p->H = bm25_heads_per_tile(p->n_tiles, k);
    p->h_tau = bm25_tau_heads(p->n_tiles, k);
    return p->H <= kBmMaxH && (int64_t)p->n_tiles * p->h_tau <= 4096;
}
// block tiles when the launch has enough of them to fill the GPU (148 SMs x 3 CTAs, twice over); quarter-block tiles
// for small corpora and single queries (more, shorter CTAs: latency)
int g_bm25_tile_chunks = 0;             // option "bm25_tile": force 1 / 2 / 4 chunks per tile (0 = automatic)
int g_bm25_inline_resolve = 1;          // option "bm25_inline_resolve": launches of <= 4 queries resolve their records in the filter kernel
int g_bm25_acc16 = 1;                   // option "bm25_acc16": 16-bit accumulators (4 CTAs per SM) in the TMA kernel
int g_bm25_by_block = 1;                // option "bm25_by_block": the TMA kernel's CTAs in (block, query) order
int g_bm25_tma = 1;                     // option "bm25_tma": block tiles through bm25_filter_tma_kernel (0: bm25_filter_kernel<4>)
static bool bm25_plan(const Bm25Device& ix, int k, int Q, Bm25Plan* p) {
    if (g_bm25_tile_chunks == 1 || g_bm25_tile_chunks == 2 || g_bm25_tile_chunks == 4)
        if (bm25_plan_for(ix, k, g_bm25_tile_chunks, p)) return true;
    if ((int64_t)Q * ix.n_blocks >= 2 * 148 * 3 && bm25_plan_for(ix, k, 4, p)) return true;
    if (bm25_plan_for(ix, k, 1, p)) return true;
    if (bm25_plan_for(ix, k, 2, p)) return true;
    return bm25_plan_for(ix, k, 4, p);
}

bool bm25_fast_supported(const Bm25Device& ix, int k, int max_query_tokens) {
    if (!ix.fast_ok || !ix.post_pack || ix.n_blocks < 1 || max_query_tokens > kBmMaxQueryTokens) return false;
    Bm25Plan a;
    return bm25_plan_for(ix, k, 1, &a) || bm25_plan_for(ix, k, 2, &a) || bm25_plan_for(ix, k, 4, &a);
}
static size_t bm25_heads_bytes(const Bm25Device& ix, int k, int Q) {
    // the larger of the two tilings (the plan of a launch depends on its query count)
    size_t per_q = 0;
    Bm25Plan p;
    for (int ch : {1, 2, 4})
        if (bm25_plan_for(ix, k, ch, &p)) per_q = std::max(per_q, (size_t)p.n_tiles * (p.H + 1) * sizeof(unsigned long long));
    return (((size_t)Q * per_q) + 255) & ~(size_t)255;
}
static int bm25_desc_stride(int max_query_tokens) { return max_query_tokens < 1 ? 1 : max_query_tokens; }
size_t bm25_fast_scratch_bytes(const Bm25Device& ix, int k, int Q, int max_query_tokens) {
    return bm25_heads_bytes(ix, k, Q) + (size_t)Q * ix.n_blocks * bm25_desc_stride(max_query_tokens) * sizeof(uint4) + 256;
}
// queries per launch: the records of a launch stay within ~256 MB of scratch (and gridDim.y within its limit)
int bm25_fast_chunk(const Bm25Device& ix, int max_query_tokens) {
    const size_t per_q = (size_t)ix.n_blocks * bm25_desc_stride(max_query_tokens) * sizeof(uint4);
    size_t n = ((size_t)256 << 20) / (per_q ? per_q : 1);
    if (n > 32768) n = 32768;
    return n < 1 ? 1 : (int)n;
}

// queries [q0, q0+Q) of the uploaded batch; outputs written at out_* + q0 (counts = -1: redo on the robust path)
cudaError_t bm25_fast_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr, int q0, int Q,
                             int max_query_tokens, const uint8_t* allow, int k, void* scratch, int32_t* out_rows,
                             double* out_scores, int32_t* out_counts, cudaStream_t st) {
    Bm25Plan pl;
    if (!bm25_plan(ix, k, Q, &pl)) return cudaErrorInvalidValue;
    const int stride = bm25_desc_stride(max_query_tokens);
    const bool use_tma = g_bm25_tma && pl.ch == 4 && stride <= kBmMaxTokens;
    unsigned long long* heads = reinterpret_cast<unsigned long long*>(scratch);
    uint4* rec = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(scratch) + bm25_heads_bytes(ix, k, Q));
    // small launches of the small-tile kernels (a single query: the latency path) resolve their records inside the
    // filter kernel: one launch less
    const bool inline_resolve = g_bm25_inline_resolve && !use_tma && pl.ch != 4 && Q <= 4;
    if (inline_resolve) {
        rec = nullptr;
    } else {
        const int64_t total = (int64_t)Q * ix.n_blocks * stride;
        int64_t g = (total + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        bm25_resolve_kernel<<<(int)g, 256, 0, st>>>(ix, d_q_terms, d_q_ptr, q0, Q, stride, use_tma ? 1 : 0, rec);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    dim3 grid_a(pl.n_tiles, Q);
    cudaError_t e = cudaSuccess;
    // 16-bit accumulators: the coarsest unit 2^cshift for which a query's tokens cannot sum to 2^16
    int cshift = 0;
    if (use_tma && g_bm25_acc16 && (size_t)ix.n_dense * ix.n_blocks * (kBmBlock / 4) < ((size_t)1 << 32)) {
        cshift = kBmDenseShift;
        while (cshift <= 8 && (int64_t)stride * ((1 << (kBmQBits - cshift)) + 1) > 65535) ++cshift;
        if (cshift > 8) cshift = 0;             // too many tokens: the unit would be too coarse to select with
    }
    if (use_tma) {
        const bool by_block = g_bm25_by_block && pl.n_tiles <= 65535;
        const dim3 grid_t = by_block ? dim3(Q, pl.n_tiles) : grid_a;
        if (cshift > 0) {
            e = ensure_dynamic_smem_of(bm25_filter_tma_kernel<true>, (size_t)(bm25_tma_smem(true)));
            if (e != cudaSuccess) return e;
            bm25_filter_tma_kernel<true><<<grid_t, kBmTmaThreads, bm25_tma_smem(true), st>>>(
                ix, rec, stride, allow, pl.H, by_block ? 1 : 0, cshift, heads);
        } else {
            e = ensure_dynamic_smem_of(bm25_filter_tma_kernel<false>, (size_t)(bm25_tma_smem(false)));
            if (e != cudaSuccess) return e;
            bm25_filter_tma_kernel<false><<<grid_t, kBmTmaThreads, bm25_tma_smem(false), st>>>(
                ix, rec, stride, allow, pl.H, by_block ? 1 : 0, 0, heads);
        }
    } else if (pl.ch == 4) {
        // (512-thread CTAs, 2 per SM, were tried for the 32 resident warps: 64 registers spill and the 16-warp
        // barriers cost more than the occupancy returns: 3.45 vs 2.61 us/query)
        e = ensure_dynamic_smem_of(bm25_filter_kernel<4, 256>, (size_t)(4 * 4096 * 4));
        if (e != cudaSuccess) return e;
        bm25_filter_kernel<4, 256><<<grid_a, 256, 4 * 4096 * 4, st>>>(ix, rec, stride, allow, pl.H, heads, d_q_terms, d_q_ptr, q0);
    } else if (pl.ch == 2) {
        bm25_filter_kernel<2, 256><<<grid_a, 256, 2 * 4096 * 4, st>>>(ix, rec, stride, allow, pl.H, heads, d_q_terms, d_q_ptr, q0);
    } else {
        bm25_filter_kernel<1, 256><<<grid_a, 256, 4096 * 4, st>>>(ix, rec, stride, allow, pl.H, heads, d_q_terms, d_q_ptr, q0);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)kBmContrib * sizeof(double) + (size_t)kBmSurvivors * (sizeof(Bm25Key) + 4);
    e = ensure_dynamic_smem_of(bm25_finish_kernel, (size_t)(smem));
    if (e != cudaSuccess) return e;
    bm25_finish_kernel<<<Q, kBmFinishThreads, smem, st>>>(ix, d_q_terms, d_q_ptr, q0, heads, pl.n_tiles, pl.H, pl.h_tau, cshift, k,
                                             out_rows + (size_t)q0 * k, out_scores + (size_t)q0 * k, out_counts + q0);
    return cudaGetLastError();
}

cudaError_t bm25_rows_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr, int Q,
                             const int32_t* d_rows, int n_rows, int k, int32_t* out_rows, double* out_scores,
                             int32_t* out_counts, cudaStream_t st) {
    int cap = 32;
    while (cap < n_rows) cap <<= 1;
    const size_t smem = (size_t)kBmContrib * sizeof(double) + (size_t)cap * sizeof(Bm25Key);
    cudaError_t e = ensure_dynamic_smem_of(bm25_rows_kernel, (size_t)(smem));
    if (e != cudaSuccess) return e;
    bm25_rows_kernel<<<Q, 256, smem, st>>>(ix, d_q_terms, d_q_ptr, d_rows, n_rows, k, out_rows, out_scores, out_counts);
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
    cudaError_t e = ensure_dynamic_smem_of(bm25_range_kernel, (size_t)(smem));
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
        e = ensure_dynamic_smem_of(bm25_select_batch_kernel, (size_t)(smem2));
        if (e != cudaSuccess) return e;
    }
    bm25_select_batch_kernel<<<Q, 256, smem2, st>>>(reinterpret_cast<const Bm25Key*>(cand), n_ranges, kp, k, out_rows,
                                                    out_scores, out_counts);
    return cudaGetLastError();
}

size_t bm25_key_bytes() { return sizeof(Bm25Key); }

}  // namespace b200rag
