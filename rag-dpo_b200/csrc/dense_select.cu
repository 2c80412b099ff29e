// dense_select.cu — the exact tail of the dense path: keep the best candidates by
// filter score, re-score those that can still reach the top-k with the CANONICAL
// fp64 dot product, order them by (score desc, row asc) and check
// the filter margin.
//
// Together with dense_scan.cu / dense_gemm.cu this replaces the selection done
// inside collection.query(...) (src/rag/retriever.py:215-220, 380-385): the
// result is the exact top-k of the fp64 brute force, ties -> lowest row.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace b200rag {

constexpr int kMergeWarps = 8;

// ---------------------------------------------------------------------------
// sample threshold: the M best of the n_keys sample keys of a query (the contraction kernel's sample pass
// leaves M per CTA), one WARP per query, registers only: every lane keeps the M best of its share, then M
// rounds of "warp-wide maximum of the lanes' heads, the owner pops".  top[q*M + M-1] is the query's threshold.
// ---------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) sample_tau_kernel(const uint64_t* __restrict__ keys, int n_queries, int n_keys,
                                                         uint64_t* __restrict__ top) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= n_queries) return;                       // warp-uniform
    const uint64_t* src = keys + (size_t)q * n_keys;
    uint64_t best[M];
#pragma unroll
    for (int i = 0; i < M; ++i) best[i] = 0ull;
    for (int i0 = 0; i0 < n_keys; i0 += 32 * 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 32 + lane;
            v[u] = i < n_keys ? src[i] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t x = v[u];
            if (x > best[M - 1]) {
#pragma unroll
                for (int i = 0; i < M; ++i) {             // sorted insert: bubble x down the list
                    const uint64_t hi = best[i] > x ? best[i] : x;
                    x = best[i] > x ? x : best[i];
                    best[i] = hi;
                }
            }
        }
    }
    for (int r = 0; r < M; ++r) {
        uint64_t mx = best[0];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const uint64_t o = __shfl_xor_sync(0xffffffffu, mx, off);
            mx = o > mx ? o : mx;
        }
        if (lane == 0) top[(size_t)q * M + r] = mx;
        if (mx != 0ull && best[0] == mx) {                // keys are distinct: exactly one owner pops
#pragma unroll
            for (int i = 0; i + 1 < M; ++i) best[i] = best[i + 1];
            best[M - 1] = 0ull;
        }
    }
}

cudaError_t sample_tau_launch(const uint64_t* keys, int n_queries, int n_keys, int m, uint64_t* top, cudaStream_t st) {
    if (m != 8) return cudaErrorInvalidValue;
    sample_tau_kernel<8><<<(n_queries + 7) / 8, 256, 0, st>>>(keys, n_queries, n_keys, top);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// canonical fp64 score, warp-cooperative.  Lane l owns the 8-element groups l, l+32, l+64,
// ... of the row — one 16-byte vector of a bf16/fp16 row, two of an fp32 row — and adds their products in
// ascending element order; then the fixed lane tree.  Valid in lane 0.
// The query sits in SHARED memory as doubles in a lane-major layout (q64_index) so that a lane's two
// doubles of every 16-byte read are conflict-free.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int q64_index(int e) {   // element e -> slot in the staged query
    const int g = e >> 3, i = e & 7;
    return ((((g >> 5) * 4 + (i >> 1)) * 32 + (g & 31)) << 1) + (i & 1);
}
__host__ __device__ __forceinline__ int q64_slots(int dim) { return ((dim / 8 + 31) / 32) * 256; }

// stages q (fp32, global) as doubles into q64s; returns |q|^2 (valid in every thread).  All threads of a
// 256-thread CTA call it; s_red: 8 doubles of shared memory.  Ends with __syncthreads().
__device__ __forceinline__ double stage_query(const float* __restrict__ q, int dim, double* q64s, double* s_red) {
    double sq = 0.0;
    for (int v = threadIdx.x; v < dim / 4; v += blockDim.x) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(q) + v);
        const int e = 4 * v;
        double2 a = make_double2((double)f.x, (double)f.y), b = make_double2((double)f.z, (double)f.w);
        *reinterpret_cast<double2*>(q64s + q64_index(e)) = a;
        *reinterpret_cast<double2*>(q64s + q64_index(e + 2)) = b;
        sq += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
    return tot;
}

template <int DT>
__device__ __forceinline__ void unpack8_f64(const uint4* v, double* x) {
    if constexpr (DT == RAG_F32) {
        x[0] = (double)__uint_as_float(v[0].x); x[1] = (double)__uint_as_float(v[0].y);
        x[2] = (double)__uint_as_float(v[0].z); x[3] = (double)__uint_as_float(v[0].w);
        x[4] = (double)__uint_as_float(v[1].x); x[5] = (double)__uint_as_float(v[1].y);
        x[6] = (double)__uint_as_float(v[1].z); x[7] = (double)__uint_as_float(v[1].w);
    } else {
        const uint32_t w[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (DT == RAG_BF16) {
                x[2 * i] = (double)__uint_as_float(w[i] << 16);
                x[2 * i + 1] = (double)__uint_as_float(w[i] & 0xFFFF0000u);
            } else {
                x[2 * i] = (double)__half2float(__ushort_as_half((unsigned short)(w[i] & 0xFFFFu)));
                x[2 * i + 1] = (double)__half2float(__ushort_as_half((unsigned short)(w[i] >> 16)));
            }
        }
    }
}

// WIDE: all loads of a 1024-d row in flight at once (small batches, registers to spare)
template <int DT, bool WIDE = false>
__device__ __forceinline__ double canonical_dot(const double* __restrict__ q64s, const void* __restrict__ rows,
                                                size_t row, int dim, int lane) {
    constexpr int VPG = DT == RAG_F32 ? 2 : 1;       // 16-byte vectors per 8-element group
    constexpr int CH = (DT == RAG_F32 && !WIDE) ? 2 : 4;   // chunks (of 32 groups) whose loads are in flight together
    const int groups = dim >> 3;
    const uint4* base = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(rows) +
                                                       row * (size_t)dim * (DT == RAG_F32 ? 4 : 2));
    double p = 0.0;
    for (int c0 = 0; c0 * 32 < groups; c0 += CH) {
        uint4 v[CH][VPG];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int g = (c0 + u) * 32 + lane;
#pragma unroll
            for (int h = 0; h < VPG; ++h) v[u][h] = g < groups ? __ldg(base + g * VPG + h) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int g = (c0 + u) * 32 + lane;
            if (g < groups) {
                double x[8];
                unpack8_f64<DT>(v[u], x);
                const double2* qq = reinterpret_cast<const double2*>(q64s) + (size_t)(c0 + u) * 4 * 32 + lane;
#pragma unroll
                for (int h = 0; h < 4; ++h) {       // exact products: one rounding per addition
                    const double2 a = qq[h * 32];
                    p = __fma_rn(a.x, x[2 * h], p);
                    p = __fma_rn(a.y, x[2 * h + 1], p);
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_down_sync(0xffffffffu, p, off));
    return p;
}

struct ExactKey {
    double s;
    uint32_t row;
    uint32_t pad;
    // "a < b" == a ranks AFTER b under (score desc, row asc)
    __device__ __forceinline__ bool operator<(const ExactKey& o) const {
        return s < o.s || (s == o.s && row > o.row);
    }
};

// ---------------------------------------------------------------------------
// fused select + refine: one CTA (8 warps) per query.
//   1. merge: every warp runs a top-kp over its share of the query's candidate keys; the 8 sorted lists
//      (even warps descending, odd ascending) are merged pairwise — elementwise max of a descending and an
//      ascending list is the bitonic top half of their union — down to ONE list of the kp best filter keys.
//   2. refine, adaptive: re-score the first m1 = k + a few candidates with the canonical fp64 dot product,
//      take the k-th exact score E_k, and re-score exactly those further candidates whose filter score can
//      still reach it (filter >= E_k - eps; the list is sorted, so they are a prefix).  Candidates past that
//      prefix provably cannot enter the top-k, so they are never fetched.
//   3. final order by rank counting over the re-scored candidates, margin check against the rows that are
//      not in the list at all (the same rigorous bound as before).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int first_round(int k) {
    const int m = (k + 6 + 7) & ~7;
    return m < 16 ? 16 : m;
}

template <int DT, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
select_refine_kernel(const uint64_t* __restrict__ cand, const int32_t* __restrict__ counts, int flat_counts, int n_lists,
                     int list_len, int sorted_lists, int32_t* __restrict__ overflow, RefineParams p) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    __shared__ double s_red[8];
    __shared__ double s_ek;
    __shared__ uint64_t s_thr;
    __shared__ int s_n;
    bool merged = false;            // block-uniform: the candidate keys are already merged in sm_keys[0..kp)
    const int kp = p.kp, b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* sm_keys = reinterpret_cast<uint64_t*>(sm_raw);                      // kMergeWarps * 2 * kp
    ExactKey* ek = reinterpret_cast<ExactKey*>(sm_raw + (size_t)kMergeWarps * 2 * kp * sizeof(uint64_t));   // kp
    double* q64s = reinterpret_cast<double*>(ek + kp);

    const double qnorm2 = stage_query(p.q + (size_t)b * p.dim, p.dim, q64s, s_red);

    // ---- 1. merge ----------------------------------------------------------------------------------
    if (cand == nullptr) {          // already merged (empty corpus): p.top holds kp keys per query
        for (int i = threadIdx.x; i < kp; i += blockDim.x) sm_keys[i] = p.top[(size_t)b * kp + i];
    } else {
        const uint64_t* src = cand + (size_t)b * n_lists * list_len;
        const int capacity = n_lists * list_len;
        WarpTopK t;
        if (flat_counts) {          // ONE contiguous list with a count (the tensor-core filter's survivors)
            const int raw = counts[b];
            const int total = raw < capacity ? raw : capacity;
            if (overflow && threadIdx.x == 0) overflow[b] = raw > capacity ? 1 : 0;
            // ---- histogram select (the usual case: ~E = 1024 survivors, kp of them wanted).  The keys stay in
            // registers (<= 8 per thread); a 256-bin histogram over [min, max] of their ordered scores finds the bin
            // that holds the kp-th best; the keys at or above that bin — kp and a few, the tail of the score
            // distribution is sparse — are compacted and ordered by rank counting.  No sort of the other ~900 keys.
            constexpr int kPerThread = 8;
            if (total <= kPerThread * (int)blockDim.x && kp >= 32) {        // block-uniform
                uint64_t key[kPerThread];
                uint32_t lo = 0xFFFFFFFFu, hi = 0u;
#pragma unroll
                for (int u = 0; u < kPerThread; ++u) {
                    const int i = (int)threadIdx.x + u * (int)blockDim.x;
                    key[u] = i < total ? src[i] : 0ull;
                    if (key[u] != 0ull) {
                        const uint32_t o = (uint32_t)(key[u] >> 32);
                        lo = o < lo ? o : lo;
                        hi = o > hi ? o : hi;
                    }
                }
                uint32_t* s_mm = reinterpret_cast<uint32_t*>(sm_keys + (size_t)8 * kp);      // [16]: per-warp min / max
                uint32_t* hist = s_mm + 16;                                                 // [256]
                uint64_t* sel = sm_keys + (size_t)2 * kp;                                   // [2 * kp] selected keys
                lo = __reduce_min_sync(0xffffffffu, lo);
                hi = __reduce_max_sync(0xffffffffu, hi);
                if (lane == 0) { s_mm[warp] = lo; s_mm[8 + warp] = hi; }
                for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
                if (threadIdx.x == 0) { s_n = 0; s_thr = 0ull; }
                __syncthreads();
#pragma unroll
                for (int w = 0; w < kMergeWarps; ++w) {
                    lo = s_mm[w] < lo ? s_mm[w] : lo;
                    hi = s_mm[8 + w] > hi ? s_mm[8 + w] : hi;
                }
                int bstar = 0;                       // keys in bins >= bstar are selected (0: all of them)
                int shift = 0;
                if (total > kp && hi > lo) {         // block-uniform
                    const uint32_t range = hi - lo;
                    shift = 32 - __clz(range) - 8;   // (range >> shift) < 256
                    if (shift < 0) shift = 0;
#pragma unroll
                    for (int u = 0; u < kPerThread; ++u)
                        if (key[u] != 0ull) atomicAdd(&hist[((uint32_t)(key[u] >> 32) - lo) >> shift], 1u);
                    __syncthreads();
                    if (warp == 0) {                 // suffix counts: lane l owns bins 8l .. 8l+7
                        uint32_t c[8], mine = 0u;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { c[j] = hist[8 * lane + j]; mine += c[j]; }
                        uint32_t suf = mine;         // inclusive suffix sum over lanes >= l
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const uint32_t v = __shfl_down_sync(0xffffffffu, suf, off);
                            if (lane + off < 32) suf += v;
                        }
                        const uint32_t above = suf - mine;           // keys in the bins of higher lanes
                        if (above < (uint32_t)kp && suf >= (uint32_t)kp) {   // exactly one lane
                            uint32_t run = above;
                            int j = 7;
                            for (; j > 0; --j) {
                                run += c[j];
                                if (run >= (uint32_t)kp) break;
                            }
                            if (j == 0) run += c[0];
                            // (the loop leaves j at the bin where the count reaches kp; run = keys at or above it)
                            s_thr = ((uint64_t)(uint32_t)(8 * lane + j) << 32) | run;
                        }
                    }
                    __syncthreads();
                    bstar = (int)(s_thr >> 32);
                }
                const int n_sel = (total > kp && hi > lo) ? (int)(uint32_t)s_thr : total;
                if (n_sel <= 2 * kp) {               // block-uniform; else (mass ties in the boundary bin) the sort below
#pragma unroll
                    for (int u = 0; u < kPerThread; ++u) {
                        if (key[u] != 0ull && (int)(((uint32_t)(key[u] >> 32) - lo) >> shift) >= bstar)
                            sel[atomicAdd(&s_n, 1)] = key[u];
                    }
                    __syncthreads();
                    const int ns = s_n;              // == n_sel, or the number of non-empty keys when all are taken
                    for (int e = threadIdx.x; e < ns; e += blockDim.x) {
                        const uint64_t me = sel[e];
                        int rank = 0;
                        for (int j = 0; j < ns; ++j) rank += sel[j] > me;          // keys are distinct
                        if (rank < kp) sm_keys[rank] = me;
                    }
                    for (int i = ns + threadIdx.x; i < kp; i += blockDim.x) sm_keys[i] = 0ull;
                    merged = true;
                }
                __syncthreads();
            }
            if (!merged) {          // 32-key blocks dealt round-robin to the warps, per-warp top-kp, pairwise merge
            t.init(sm_keys + (size_t)warp * 2 * kp, kp, lane);
            for (int i0 = warp * 32; i0 < total; i0 += kMergeWarps * 32) {
                const int i = i0 + lane;
                t.offer(i < total ? src[i] : 0ull, lane);
            }
            }
        } else {                    // n_lists lists of list_len keys (0 = empty): warp w takes lists w, w+8, ...
            const int32_t* cnt = counts ? counts + (size_t)b * n_lists : nullptr;
            // Sorted lists (the scan kernel's per-CTA top-kp lists): the kp-th largest list HEAD is a lower bound
            // of the kp-th best key overall (kp distinct rows reach it), so only the list prefixes >= that bound
            // can matter: kp + a handful of keys instead of n_lists * kp.  They are collected and sorted directly.
            if (sorted_lists && !cnt && n_lists >= kp && n_lists <= (int)blockDim.x) {
                const int cap = kMergeWarps * 2 * kp;
                const uint64_t head = (int)threadIdx.x < n_lists ? src[(size_t)threadIdx.x * list_len] : 0ull;
                sm_keys[threadIdx.x] = head;
                if (threadIdx.x == 0) { s_thr = 0ull; s_n = 0; }
                __syncthreads();
                if (head != 0ull) {
                    int rank = 0;
                    for (int j = 0; j < n_lists; ++j) rank += sm_keys[j] > head;
                    if (rank == kp - 1) s_thr = head;          // keys are distinct: at most one head has this rank
                }
                __syncthreads();
                const uint64_t thr = s_thr;
                if (thr != 0ull) {                             // block-uniform
                    if ((int)threadIdx.x < n_lists) {
                        const uint64_t* lp = src + (size_t)threadIdx.x * list_len;
                        for (int i = 0; i < list_len; ++i) {
                            const uint64_t e = lp[i];
                            if (e < thr) break;                // also stops at the empty tail (key 0)
                            const int slot = atomicAdd(&s_n, 1);
                            if (slot < cap) sm_keys[slot] = e;
                        }
                    }
                    __syncthreads();
                    const int n = s_n;
                    if (n <= cap) {                            // block-uniform
                        int P = kp;
                        while (P < n) P <<= 1;
                        for (int i = n + threadIdx.x; i < P; i += blockDim.x) sm_keys[i] = 0ull;
                        block_bitonic_desc(sm_keys, P);
                        merged = true;
                    }
                }
                __syncthreads();
            }
            if (!merged) {
            t.init(sm_keys + (size_t)warp * 2 * kp, kp, lane);      // (a block barrier separates this from the above)
            if (overflow && counts) {
                int over = 0;
                for (int l = threadIdx.x; l < n_lists; l += blockDim.x) over |= cnt[l] > list_len;
                over = __syncthreads_or(over);
                if (threadIdx.x == 0) overflow[b] = over ? 1 : 0;
            }
            for (int l = warp; l < n_lists; l += kMergeWarps) {
                const int n = cnt ? min(cnt[l], list_len) : list_len;
                const uint64_t* lp = src + (size_t)l * list_len;
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    t.offer(i < n ? lp[i] : 0ull, lane);
                }
            }
            }
        }
        if (!merged) {
        t.finish(lane);             // buf[0..n) sorted descending, the rest empty
        if (warp & 1) {             // odd warps: ascending, so that (even, odd) pairs are bitonic
            for (int j = lane; j < kp / 2; j += 32) {
                const uint64_t a = t.buf[j], c = t.buf[kp - 1 - j];
                t.buf[j] = c; t.buf[kp - 1 - j] = a;
            }
        }
        // pairwise reduction 8 -> 4 -> 2 -> 1 lists; list i of a round lives at slot i * span
        for (int span = 1; span < kMergeWarps; span <<= 1) {
            const int npairs = kMergeWarps / (2 * span);
            __syncthreads();
            for (int tt = threadIdx.x; tt < kp * npairs; tt += blockDim.x) {
                const int pr = tt / kp, j = tt - pr * kp;
                uint64_t* A = sm_keys + (size_t)pr * 2 * span * 2 * kp;
                const uint64_t c = A[(size_t)span * 2 * kp + j];
                if (c > A[j]) A[j] = c;                      // top half of (descending A, ascending B): bitonic
            }
            for (int stride = kp >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int tt = threadIdx.x; tt < (kp >> 1) * npairs; tt += blockDim.x) {
                    const int pr = tt / (kp >> 1), u = tt - pr * (kp >> 1);
                    uint64_t* A = sm_keys + (size_t)pr * 2 * span * 2 * kp;
                    const int lo = 2 * u - (u & (stride - 1)), hi = lo + stride;
                    const uint64_t x = A[lo], y = A[hi];
                    const bool swap = (pr & 1) ? (y < x) : (x < y);     // list pr of the next round: even -> descending
                    if (swap) { A[lo] = y; A[hi] = x; }
                }
            }
        }
        }
    }
    __syncthreads();
    const uint64_t* top = sm_keys;                  // kp best filter keys, descending, empties (0) last

    // ---- 2. adaptive refine ------------------------------------------------------------------------
    int nc = 0;                                     // candidates in the list
    for (int i0 = 0; i0 < kp; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        nc += __syncthreads_count(i < kp && top[i] != 0ull);
    }
    const double qn = sqrt(qnorm2);
    double eps = (p.eps_rel * qn + (p.q_resid ? (double)p.q_resid[b] : 0.0)) * (double)(*p.max_row_norm);
    if (p.x_resid) eps += 1.004 * qn * (double)(*p.x_resid);

    const int m1 = min(nc, first_round(p.k));
    for (int j = warp; j < m1; j += kMergeWarps) {
        const uint32_t row = key_row(top[j]);
        const double s = canonical_dot<DT, MIN_BLOCKS <= 2>(q64s, p.rows, row, p.dim, lane);
        if (lane == 0) { ek[j].s = s; ek[j].row = row; ek[j].pad = 0; }
    }
    __syncthreads();
    int m2 = m1;
    if (m1 >= p.k && m1 < nc) {                     // block-uniform
        if (threadIdx.x < m1) {
            const ExactKey me = ek[threadIdx.x];
            int rank = 0;
            for (int i = 0; i < m1; ++i) rank += me < ek[i];
            if (rank == p.k - 1) s_ek = me.s;
        }
        __syncthreads();
        const double e_k = s_ek;
        // candidates whose filter score can still reach the k-th exact score: a prefix of the sorted list
        for (int i0 = m1; i0 < nc; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            const int need = __syncthreads_count(i < nc && !(e_k > (double)key_score(top[i]) + eps));
            m2 += need;
            if (need < (int)blockDim.x) break;      // block-uniform
        }
        for (int j = m1 + warp; j < m2; j += kMergeWarps) {
            const uint32_t row = key_row(top[j]);
            const double s = canonical_dot<DT, MIN_BLOCKS <= 2>(q64s, p.rows, row, p.dim, lane);
            if (lane == 0) { ek[j].s = s; ek[j].row = row; ek[j].pad = 0; }
        }
        __syncthreads();
    }

    // ---- 3. final order (rank counting), outputs, margin check -------------------------------------
    const int nout = m2 < p.k ? m2 : p.k;
    for (int j = threadIdx.x; j < m2; j += blockDim.x) {
        const ExactKey me = ek[j];
        int rank = 0;
        for (int i = 0; i < m2; ++i) rank += me < ek[i];
        if (rank < p.k) {
            if (p.out_gids) p.out_gids[(size_t)b * p.k + rank] = shard_global_row(me.row, p.shard, p.n_shards, p.shard_block);
            else p.out_rows[(size_t)b * p.k + rank] = (int32_t)me.row;
            p.out_scores[(size_t)b * p.k + rank] = me.s;
            if (rank == p.k - 1) s_ek = me.s;
        }
    }
    for (int i = nout + threadIdx.x; i < p.k; i += blockDim.x) {
        if (p.out_gids) p.out_gids[(size_t)b * p.k + i] = -1;
        else p.out_rows[(size_t)b * p.k + i] = -1;
        p.out_scores[(size_t)b * p.k + i] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.out_counts[b] = nout;
        int flag = 0;
        float tau = 0.f;
        const bool have_k = m2 >= p.k;
        const double e_k = have_k ? s_ek : 0.0;
        // upper bound of the filter score of every row that was NOT re-scored
        bool bounded = false;
        double outside = 0.0;
        if (m2 < nc) {                              // next candidate of the sorted list (safe by construction)
            bounded = true;
            outside = (double)key_score(top[m2]);
        } else if (nc == kp) {                      // list full: its last key bounds everything that was dropped
            bounded = true;
            outside = (double)key_score(top[kp - 1]);
        } else if (p.tau_keys) {                    // threshold capture: everything >= tau_q is in the list
            const uint64_t tk = p.tau_keys[(size_t)b * p.tau_stride + p.tau_stride - 1];
            if (tk != 0ull) { bounded = true; outside = (double)key_score(tk); }
        }
        const bool over = overflow && cand && overflow[b] != 0;
        if (bounded || over) {
            // safe iff the k-th exact score beats (outside + eps); an overflowed list voids the bound
            if (over || !have_k || !(e_k > outside + eps)) {
                flag = 1;
                tau = have_k ? __double2float_rd(e_k - eps) : -3.0e38f;
                atomicAdd(p.n_flagged, 1);
            }
        }
        p.flags[b] = flag;
        p.tau[b] = tau;
    }
    // ---- 4. device-driven fallback: the last CTA of the grid compacts the flagged queries --------------------
    if (p.fb_done != nullptr) {
        __shared__ int s_last, s_base, s_wsum[8];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();                                   // this CTA's flag / tau are visible device-wide
            s_last = atomicAdd(p.fb_done, 1u) == gridDim.x - 1 ? 1 : 0;
            s_base = 0;
        }
        __syncthreads();
        if (s_last) {                                          // block-uniform
            __threadfence();
            for (int b0 = 0; b0 < p.B; b0 += blockDim.x) {
                const int bb = b0 + threadIdx.x;
                const bool f = bb < p.B && reinterpret_cast<volatile int32_t*>(p.flags)[bb] != 0;
                const unsigned m = __ballot_sync(0xffffffffu, f);
                if (lane == 0) s_wsum[warp] = __popc(m);
                __syncthreads();
                int off = s_base;
                for (int w = 0; w < warp; ++w) off += s_wsum[w];
                const int slot = off + __popc(m & ((1u << lane) - 1u));
                if (f) {
                    if (slot < p.fb_max) {
                        p.fb_index[slot] = bb;
                        p.fb_tau[slot] = reinterpret_cast<volatile float*>(p.tau)[bb];
                    } else {
                        p.out_counts[bb] = -1;                 // more fallbacks than one stream-ordered call serves
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int tot = 0;
                    for (int w = 0; w < 8; ++w) tot += s_wsum[w];
                    s_base += tot;
                }
                __syncthreads();
            }
            const int nfb = s_base < p.fb_max ? s_base : p.fb_max;
            for (int s = 0; s < nfb; ++s) {
                const int qb = p.fb_index[s];
                for (int c = threadIdx.x; c < p.dim; c += blockDim.x) p.fb_q[(size_t)s * p.dim + c] = p.q[(size_t)qb * p.dim + c];
            }
            if ((int)threadIdx.x < p.fb_max) p.fb_counts[threadIdx.x] = 0u;
            if (threadIdx.x == 0) { *p.fb_n = nfb; *p.fb_done = 0u; }
        }
    }
}

static size_t select_refine_smem(int kp, int dim) {
    return (size_t)kMergeWarps * 2 * kp * sizeof(uint64_t) + (size_t)kp * sizeof(ExactKey) +
           (size_t)q64_slots(dim) * sizeof(double);
}

template <int DT>
static cudaError_t select_refine_dt(const uint64_t* cand, const int32_t* counts, int flat_counts, int n_lists,
                                    int list_len, int sorted_lists, int32_t* overflow, const RefineParams& p,
                                    cudaStream_t st) {
    const size_t smem = select_refine_smem(p.kp, p.dim);
    // big batches: 6 (fp32 rows: 5, the registers of the wider loads) CTAs per SM; MIN_BLOCKS = 2 keeps the
    // registers that let the refine loads overlap, which is what matters for a handful of queries
    constexpr int kBig = DT == RAG_F32 ? 5 : 6;
    auto kern = p.B >= 512 ? select_refine_kernel<DT, kBig> : select_refine_kernel<DT, 2>;
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_dynamic_smem_of(kern, (size_t)(smem));
        if (e != cudaSuccess) return e;
    }
    kern<<<p.B, 256, smem, st>>>(cand, counts, flat_counts, n_lists, list_len, sorted_lists, overflow, p);
    return cudaGetLastError();
}

cudaError_t merge_refine_launch(const uint64_t* cand, const int32_t* counts, int flat_counts, int n_lists, int list_len,
                                int sorted_lists, int32_t* overflow, const RefineParams& p, cudaStream_t st) {
    if (p.dtype == RAG_F32)
        return select_refine_dt<RAG_F32>(cand, counts, flat_counts, n_lists, list_len, sorted_lists, overflow, p, st);
    if (p.dtype == RAG_BF16)
        return select_refine_dt<RAG_BF16>(cand, counts, flat_counts, n_lists, list_len, sorted_lists, overflow, p, st);
    return select_refine_dt<RAG_F16>(cand, counts, flat_counts, n_lists, list_len, sorted_lists, overflow, p, st);
}

// p.top already holds the merged kp keys per query (used for an empty corpus: all keys 0)
cudaError_t refine_launch(const RefineParams& p, cudaStream_t st) {
    return merge_refine_launch(nullptr, nullptr, 0, 0, 0, 0, nullptr, p, st);
}

// ---------------------------------------------------------------------------
// fallback tail: exact scores of every collected row, then a streaming
// best-k select (sort 2048 at a time, keep the head).
// ---------------------------------------------------------------------------
constexpr int kSelN = 2048;

template <int DT>
__global__ void __launch_bounds__(256) collect_select_kernel(CollectSelectParams p) {
    __shared__ ExactKey ek[kSelN];
    __shared__ double s_red[8];
    extern __shared__ __align__(16) uint8_t sm_raw[];
    double* q64s = reinterpret_cast<double*>(sm_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x;
    if (p.n_active_dev != nullptr && qi >= *p.n_active_dev) return;      // block-uniform
    if (p.counts[qi] > (unsigned)p.cap) {           // more rows at the bound than the collect list holds: the
        if (threadIdx.x == 0) p.out_counts[p.query_index[qi]] = -1;      // host-checked call redoes this query
        return;
    }
    const uint32_t* list = p.rows_list + (size_t)qi * p.cap;
    double* sc = p.scratch_scores + (size_t)qi * p.cap;
    const int cnt = (int)(p.counts[qi] < (unsigned)p.cap ? p.counts[qi] : (unsigned)p.cap);
    stage_query(p.q + (size_t)qi * p.dim, p.dim, q64s, s_red);

    for (int j = warp; j < cnt; j += 8) {
        double s = canonical_dot<DT>(q64s, p.rows, list[j], p.dim, lane);
        if (lane == 0) sc[j] = s;
    }
    __syncthreads();
    const int keep = p.k;                       // head kept between rounds
    for (int i = threadIdx.x; i < kSelN; i += blockDim.x) { ek[i].s = -INFINITY; ek[i].row = 0xFFFFFFFFu; ek[i].pad = 0; }
    __syncthreads();
    const int chunk = kSelN - keep;
    for (int base = 0; base < cnt || base == 0; base += chunk) {
        for (int i = threadIdx.x; i < chunk; i += blockDim.x) {
            int j = base + i;
            ExactKey e;
            e.pad = 0;
            if (j < cnt) { e.s = sc[j]; e.row = list[j]; } else { e.s = -INFINITY; e.row = 0xFFFFFFFFu; }
            ek[keep + i] = e;
        }
        block_bitonic_desc(ek, kSelN);
        if (cnt == 0) break;
    }
    const int out = p.query_index[qi];
    int nout = cnt < p.k ? cnt : p.k;
    for (int i = threadIdx.x; i < p.k; i += blockDim.x) {
        if (p.out_gids)
            p.out_gids[(size_t)out * p.k + i] = i < nout ? shard_global_row(ek[i].row, p.shard, p.n_shards, p.shard_block) : -1;
        else
            p.out_rows[(size_t)out * p.k + i] = i < nout ? (int32_t)ek[i].row : -1;
        p.out_scores[(size_t)out * p.k + i] = i < nout ? ek[i].s : 0.0;
    }
    if (threadIdx.x == 0) p.out_counts[out] = nout;
}

cudaError_t collect_select_launch(const CollectSelectParams& p, cudaStream_t st) {
    const size_t smem = (size_t)q64_slots(p.dim) * sizeof(double);
    if (p.dtype == RAG_F32) collect_select_kernel<RAG_F32><<<p.nq, 256, smem, st>>>(p);
    else if (p.dtype == RAG_BF16) collect_select_kernel<RAG_BF16><<<p.nq, 256, smem, st>>>(p);
    else collect_select_kernel<RAG_F16><<<p.nq, 256, smem, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// multi-GPU exchange tail: G ranks' exact (score, global id) lists -> top-k
// ---------------------------------------------------------------------------
struct ExactKey64 {
    double s;
    int64_t id;
    __device__ __forceinline__ bool operator<(const ExactKey64& o) const {
        return s < o.s || (s == o.s && id > o.id);
    }
};

__global__ void __launch_bounds__(256) merge_exact_kernel(const double* __restrict__ scores,
                                                          const int64_t* __restrict__ ids, int G, int B, int k,
                                                          int64_t rank_stride, int nsort, double* out_scores,
                                                          int64_t* out_ids, int32_t* out_counts) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    ExactKey64* ek = reinterpret_cast<ExactKey64*>(sm_raw);
    __shared__ int s_count;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < nsort; i += blockDim.x) {
        ExactKey64 e;
        e.s = -INFINITY; e.id = INT64_MAX;
        if (i < G * k) {
            int g = i / k, j = i % k;
            size_t src = (size_t)g * rank_stride + (size_t)b * k + j;
            int64_t id = ids[src];
            if (id >= 0) { e.s = scores[src]; e.id = id; ++local; }
        }
        ek[i] = e;
    }
    atomicAdd(&s_count, local);
    block_bitonic_desc(ek, nsort);
    const int nout = s_count < k ? s_count : k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out_ids[(size_t)b * k + i] = i < nout ? ek[i].id : -1;
        out_scores[(size_t)b * k + i] = i < nout ? ek[i].s : 0.0;
    }
    if (threadIdx.x == 0) out_counts[b] = nout;
}

cudaError_t merge_exact_launch(const double* scores, const int64_t* ids, int G, int B, int k, int64_t rank_stride,
                               double* out_scores, int64_t* out_ids, int32_t* out_counts, cudaStream_t st) {
    int nsort = 32;
    while (nsort < G * k) nsort <<= 1;
    size_t smem = (size_t)nsort * sizeof(ExactKey64);
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_dynamic_smem_of(merge_exact_kernel, (size_t)(smem));
        if (e != cudaSuccess) return e;
    }
    merge_exact_kernel<<<B, 256, smem, st>>>(scores, ids, G, B, k, rank_stride, nsort, out_scores, out_ids, out_counts);
    return cudaGetLastError();
}

}  // namespace b200rag
