// dense_select.cu — the exact tail of the dense path: merge the per-warp
// candidate lists, re-score the survivors with the CANONICAL fp64 dot product
// (DESIGN.md §3), sort by (score desc, row asc) and check the filter margin.
//
// Together with dense_scan.cu / dense_gemm.cu this replaces the selection done
// inside collection.query(...) (src/rag/retriever.py:215-220, 380-385): the
// result is the exact top-k of the fp64 brute force, ties -> lowest row.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace b200rag {

// ---------------------------------------------------------------------------
// merge: one CTA per query, 8 warps each scanning a strided share of the keys
// ---------------------------------------------------------------------------
constexpr int kMergeWarps = 8;

// leaves the query's kp best keys, sorted descending, in sm_keys[0..kp) (all threads synchronised)
__device__ __forceinline__ void merge_body(const uint64_t* __restrict__ cand, const int32_t* __restrict__ counts,
                                           int flat_counts, int n_lists, int list_len, int kp, int b,
                                           uint64_t* sm_keys, int32_t* __restrict__ overflow) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t* src = cand + (size_t)b * n_lists * list_len;
    const int32_t* cnt = (counts && !flat_counts) ? counts + (size_t)b * n_lists : nullptr;
    const int flat_total = (counts && flat_counts) ? counts[b] : 0;

    if (overflow && counts) {
        int over = 0;
        if (flat_counts) {
            over = flat_total > n_lists * list_len;
        } else {
            for (int l = threadIdx.x; l < n_lists; l += blockDim.x) over |= cnt[l] > list_len;
            over = __syncthreads_or(over);
        }
        if (threadIdx.x == 0) overflow[b] = over ? 1 : 0;
    }
    WarpTopK t;
    t.init(sm_keys + (size_t)warp * 2 * kp, kp, lane);
    // each warp walks whole lists: warp w takes lists w, w+8, ...
    for (int l = warp; l < n_lists; l += kMergeWarps) {
        int n = list_len;
        if (cnt) n = min(cnt[l], list_len);
        else if (flat_counts) n = max(0, min(flat_total - l * list_len, list_len));
        const uint64_t* lp = src + (size_t)l * list_len;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            t.offer(i < n ? lp[i] : 0ull, lane);
        }
    }
    t.finish(lane);
    __syncthreads();
    block_bitonic_desc(sm_keys, kMergeWarps * 2 * kp);      // power of two
}

__global__ void __launch_bounds__(kMergeWarps * 32)
merge_kernel(const uint64_t* __restrict__ cand, const int32_t* __restrict__ counts, int flat_counts, int n_lists,
             int list_len, int kp, uint64_t* __restrict__ top, int32_t* __restrict__ overflow) {
    extern __shared__ __align__(16) uint64_t sm_keys[];   // kMergeWarps * 2 * kp
    const int b = blockIdx.x;
    merge_body(cand, counts, flat_counts, n_lists, list_len, kp, b, sm_keys, overflow);
    for (int i = threadIdx.x; i < kp; i += blockDim.x) top[(size_t)b * kp + i] = sm_keys[i];
}

cudaError_t merge_launch(const uint64_t* cand, const int32_t* counts, int flat_counts, int B, int n_lists, int list_len,
                         int kp, uint64_t* top, int32_t* overflow, cudaStream_t st) {
    size_t smem = (size_t)kMergeWarps * 2 * kp * sizeof(uint64_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    merge_kernel<<<B, kMergeWarps * 32, smem, st>>>(cand, counts, flat_counts, n_lists, list_len, kp, top, overflow);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// canonical fp64 score (warp-cooperative).  Valid in lane 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double load_elem_f64(const void* rows, int dtype, size_t idx) {
    if (dtype == RAG_F32) return (double)reinterpret_cast<const float*>(rows)[idx];
    if (dtype == RAG_BF16) return (double)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[idx]);
    return (double)__half2float(reinterpret_cast<const __half*>(rows)[idx]);
}

__device__ __forceinline__ double canonical_dot(const float* __restrict__ q, const void* __restrict__ rows, int dtype,
                                                size_t row, int dim, int lane) {
    double p = 0.0;
    const size_t base = row * (size_t)dim;
    const int steps = dim / 32;
    // the summation ORDER is fixed (j ascending); the loads are issued 8 at a time so that their latencies overlap
    int j = 0;
    for (; j + 8 <= steps; j += 8) {
        double a[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a[u] = (double)q[32 * (j + u) + lane];
            x[u] = load_elem_f64(rows, dtype, base + 32 * (j + u) + lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) p = __fma_rn(a[u], x[u], p);   // exact product: one rounding, p + a*x
    }
    for (; j < steps; ++j) {
        const double a = (double)q[32 * j + lane];
        const double x = load_elem_f64(rows, dtype, base + 32 * j + lane);
        p = __fma_rn(a, x, p);
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

// canonical re-score of the kp candidates in `top` (shared or global memory), final order, margin check
__device__ __forceinline__ void refine_body(const RefineParams& p, int b, const uint64_t* top, ExactKey* ek) {
    __shared__ double s_qnorm2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* q = p.q + (size_t)b * p.dim;
    const int nsort = p.kp < 32 ? 32 : p.kp;

    for (int j = warp; j < nsort; j += 8) {
        uint64_t key = j < p.kp ? top[j] : 0ull;
        ExactKey e;
        e.pad = 0;
        if (key == 0ull) {
            e.s = -INFINITY; e.row = 0xFFFFFFFFu;
        } else {
            e.row = key_row(key);
            e.s = canonical_dot(q, p.rows, p.dtype, e.row, p.dim, lane);
        }
        if (lane == 0) ek[j] = e;
    }
    if (warp == 0) {
        double s = 0.0;
        for (int i = lane; i < p.dim; i += 32) s += (double)q[i] * (double)q[i];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) s_qnorm2 = s;
    }
    __syncthreads();
    if (warp != 0) return;
    warp_bitonic_desc(ek, nsort, lane);

    int count = 0;
    for (int i = lane; i < p.kp; i += 32) count += (ek[i].row != 0xFFFFFFFFu);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
    const int nout = count < p.k ? count : p.k;
    for (int i = lane; i < p.k; i += 32) {
        p.out_rows[(size_t)b * p.k + i] = i < nout ? (int32_t)ek[i].row : -1;
        p.out_scores[(size_t)b * p.k + i] = i < nout ? ek[i].s : 0.0;
    }
    if (lane == 0) {
        p.out_counts[b] = nout;
        int flag = 0;
        float tau = 0.f;
        const double qn = sqrt(s_qnorm2);
        double eps = (p.eps_rel * qn + (p.q_resid ? (double)p.q_resid[b] : 0.0)) * (double)(*p.max_row_norm);
        if (p.x_resid) eps += 1.004 * qn * (double)(*p.x_resid);
        // upper bound of the filter score of every row that is NOT a candidate
        bool bounded = false;
        double outside = 0.0;
        if (count == p.kp) {                       // list full: the kp-th candidate bounds the rest
            bounded = true;
            outside = (double)key_score(top[p.kp - 1]);
        } else if (p.tau_keys) {                   // threshold capture: everything >= tau_q was kept
            const uint64_t tk = p.tau_keys[(size_t)b * p.tau_stride + p.tau_stride - 1];
            if (tk != 0ull) { bounded = true; outside = (double)key_score(tk); }
        }
        const bool over = p.overflow && p.overflow[b] != 0;
        if (bounded || over) {
            // safe iff the k-th exact score beats (outside + eps); an overflowed list voids the bound
            const bool have_k = count >= p.k;
            const double e_k = have_k ? ek[p.k - 1].s : 0.0;
            if (over || !have_k || !(e_k > outside + eps)) {
                flag = 1;
                tau = have_k ? __double2float_rd(e_k - eps) : -3.0e38f;
                atomicAdd(p.n_flagged, 1);
            }
        }
        p.flags[b] = flag;
        p.tau[b] = tau;
    }
}

__global__ void __launch_bounds__(256) refine_kernel(RefineParams p) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    refine_body(p, blockIdx.x, p.top + (size_t)blockIdx.x * p.kp, reinterpret_cast<ExactKey*>(sm_raw));
}

cudaError_t refine_launch(const RefineParams& p, cudaStream_t st) {
    int nsort = p.kp < 32 ? 32 : p.kp;
    refine_kernel<<<p.B, 256, (size_t)nsort * sizeof(ExactKey), st>>>(p);
    return cudaGetLastError();
}

// merge + refine in one launch: the merged candidates never leave shared memory
// MIN_BLOCKS = 7 (32 registers): a 1024-query batch is a single wave on 148 SMs; MIN_BLOCKS = 2 keeps the
// registers that let the refine loads overlap, which is what matters for a handful of queries
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
merge_refine_kernel(const uint64_t* __restrict__ cand, const int32_t* __restrict__ counts, int flat_counts, int n_lists,
                    int list_len, int32_t* __restrict__ overflow, RefineParams p) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    uint64_t* sm_keys = reinterpret_cast<uint64_t*>(sm_raw);                        // kMergeWarps * 2 * kp
    ExactKey* ek = reinterpret_cast<ExactKey*>(sm_raw + (size_t)kMergeWarps * 2 * p.kp * sizeof(uint64_t));
    merge_body(cand, counts, flat_counts, n_lists, list_len, p.kp, blockIdx.x, sm_keys, overflow);
    refine_body(p, blockIdx.x, sm_keys, ek);
}

cudaError_t merge_refine_launch(const uint64_t* cand, const int32_t* counts, int flat_counts, int n_lists, int list_len,
                                int32_t* overflow, const RefineParams& p, cudaStream_t st) {
    const int nsort = p.kp < 32 ? 32 : p.kp;
    const size_t smem = (size_t)kMergeWarps * 2 * p.kp * sizeof(uint64_t) + (size_t)nsort * sizeof(ExactKey);
    auto kern = p.B >= 512 ? merge_refine_kernel<7> : merge_refine_kernel<2>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<p.B, 256, smem, st>>>(cand, counts, flat_counts, n_lists, list_len, overflow, p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// fallback tail: exact scores of every collected row, then a streaming
// best-k select (sort 2048 at a time, keep the head).
// ---------------------------------------------------------------------------
constexpr int kSelN = 2048;

__global__ void __launch_bounds__(256) collect_select_kernel(CollectSelectParams p) {
    __shared__ ExactKey ek[kSelN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x;
    const float* q = p.q + (size_t)qi * p.dim;
    const uint32_t* list = p.rows_list + (size_t)qi * p.cap;
    double* sc = p.scratch_scores + (size_t)qi * p.cap;
    const int cnt = (int)(p.counts[qi] < (unsigned)p.cap ? p.counts[qi] : (unsigned)p.cap);

    for (int j = warp; j < cnt; j += 8) {
        double s = canonical_dot(q, p.rows, p.dtype, list[j], p.dim, lane);
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
        p.out_rows[(size_t)out * p.k + i] = i < nout ? (int32_t)ek[i].row : -1;
        p.out_scores[(size_t)out * p.k + i] = i < nout ? ek[i].s : 0.0;
    }
    if (threadIdx.x == 0) p.out_counts[out] = nout;
}

cudaError_t collect_select_launch(const CollectSelectParams& p, cudaStream_t st) {
    collect_select_kernel<<<p.nq, 256, 0, st>>>(p);
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
        cudaError_t e = cudaFuncSetAttribute(merge_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    merge_exact_kernel<<<B, 256, smem, st>>>(scores, ids, G, B, k, rank_stride, nsort, out_scores, out_ids, out_counts);
    return cudaGetLastError();
}

}  // namespace b200rag
