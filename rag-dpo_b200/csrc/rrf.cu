// rrf.cu — weighted Reciprocal Rank Fusion for a batch of questions.
//
// Replaces reciprocal_rank_fusion (src/rag/retriever.py:66-90) and the fusion
// tail (src/rag/retriever.py:454-467): scores[id] += w_r / (k + rank + 1)
// accumulated in ranking order (fp64, correctly rounded divide and add), then a
// STABLE descending sort over first-seen order and a cut to `top`.
//
// One CTA per question; R*L <= 8192 entries live in (dynamic) shared memory.  The work
// per question is ~400 entries, so the kernel is latency-bound by design; it
// exists so that fused lists never leave the device when many questions are
// served per call.
#include "common.cuh"
#include "kernels.h"

namespace b200rag {

constexpr int kRrfMaxEntries = 8192;

__global__ void __launch_bounds__(256)
rrf_kernel(const int32_t* __restrict__ ids, const double* __restrict__ weights, int R, int L, int rrf_k, int top,
           int32_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) uint8_t rrf_smem[];
    __shared__ int s_distinct;
    const int q = blockIdx.x;
    const int n = R * L;
    const int np = (n + 7) & ~7;
    double* s_score = reinterpret_cast<double*>(rrf_smem);              // np
    int32_t* s_id = reinterpret_cast<int32_t*>(s_score + np);           // np
    int16_t* s_rank = reinterpret_cast<int16_t*>(s_id + np);            // np: enumerate() index inside its ranking
    uint8_t* s_first = reinterpret_cast<uint8_t*>(s_rank + np);         // np: 1 = first occurrence of its id
    const int32_t* my_ids = ids + (size_t)q * n;
    const double* w = weights + (size_t)q * R;

    if (threadIdx.x == 0) s_distinct = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) s_id[e] = my_ids[e];
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int r = e / L, j = e % L;
        int rank = 0;
        for (int jj = 0; jj < j; ++jj) rank += (s_id[r * L + jj] >= 0);
        s_rank[e] = (int16_t)rank;
        const int32_t id = s_id[e];
        bool first = id >= 0;
        for (int ee = 0; first && ee < e; ++ee) first = (s_id[ee] != id);
        s_first[e] = first ? 1 : 0;
    }
    __syncthreads();
    int local = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        if (!s_first[e]) continue;
        ++local;
        const int32_t id = s_id[e];
        double acc = 0.0;
        for (int ee = e; ee < n; ++ee) {               // later occurrences, in ranking order
            if (s_id[ee] == id)
                acc = __dadd_rn(acc, __ddiv_rn(w[ee / L], (double)(rrf_k + (int)s_rank[ee] + 1)));
        }
        s_score[e] = acc;
    }
    if (local) atomicAdd(&s_distinct, local);
    __syncthreads();
    // stable descending order by counting: position = #entries that rank before
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        if (!s_first[e]) continue;
        const double sc = s_score[e];
        int pos = 0;
        for (int ee = 0; ee < n; ++ee) {
            if (!s_first[ee] || ee == e) continue;
            const double o = s_score[ee];
            pos += (o > sc) || (o == sc && ee < e);
        }
        if (pos < top) {
            out_ids[(size_t)q * top + pos] = s_id[e];
            out_scores[(size_t)q * top + pos] = sc;
        }
    }
    const int nout = s_distinct < top ? s_distinct : top;
    for (int i = nout + threadIdx.x; i < top; i += blockDim.x) {
        out_ids[(size_t)q * top + i] = -1;
        out_scores[(size_t)q * top + i] = 0.0;
    }
    if (threadIdx.x == 0) out_counts[q] = nout;
}

// ---------------------------------------------------------------------------
// Rerank select: the output step of CrossEncoderReranker.rerank (src/rag/reranker.py:172-211) for a batch of
// questions.  final = float(score) + topic boost (fp64 add), STABLE descending order (list.sort(reverse=True) keeps
// the original order of equal scores), cut to top_k, drop entries below min_score, but never return fewer than 3
// when at least 3 candidates exist.  One CTA per question, L <= 1024 candidates in shared memory.
// ---------------------------------------------------------------------------
constexpr int kRerankMax = 1024;

__global__ void __launch_bounds__(256)
rerank_select_kernel(const float* __restrict__ scores, const double* __restrict__ boosts, const int32_t* __restrict__ lens,
                     int L, int top_k, double min_score, int32_t* __restrict__ out_idx, double* __restrict__ out_scores,
                     int32_t* __restrict__ out_counts) {
    __shared__ double s_final[kRerankMax];
    __shared__ int s_kept;
    const int q = blockIdx.x;
    const int n = lens ? (lens[q] < L ? lens[q] : L) : L;
    if (threadIdx.x == 0) s_kept = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double sc = (double)scores[(size_t)q * L + i];
        s_final[i] = boosts ? __dadd_rn(sc, boosts[(size_t)q * L + i]) : sc;
    }
    __syncthreads();
    const int cut = top_k < n ? top_k : n;
    int local = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double sc = s_final[i];
        int pos = 0;
        for (int j = 0; j < n; ++j) {
            const double o = s_final[j];
            pos += (o > sc) || (o == sc && j < i);
        }
        if (pos < cut) {
            out_idx[(size_t)q * top_k + pos] = i;
            out_scores[(size_t)q * top_k + pos] = sc;
            local += sc >= min_score;
        }
    }
    if (local) atomicAdd(&s_kept, local);
    __syncthreads();
    // the entries >= min_score of a descending list are a prefix; fewer than 3 of them: the first 3 ranked instead
    int kept = s_kept;
    if (kept < 3 && n >= 3) kept = 3 < cut ? 3 : cut;
    // (top_k < 3: the reference's ranked[:3] would return 3; the output holds top_k slots, so such calls are refused
    // by the entry point)
    __syncthreads();
    for (int i = kept + threadIdx.x; i < top_k; i += blockDim.x) {
        out_idx[(size_t)q * top_k + i] = -1;
        out_scores[(size_t)q * top_k + i] = 0.0;
    }
    if (threadIdx.x == 0) out_counts[q] = kept;
}

int rerank_max_candidates() { return kRerankMax; }

cudaError_t rerank_select_launch(const float* scores, const double* boosts, const int32_t* lens, int Q, int L, int top_k,
                                 double min_score, int32_t* out_idx, double* out_scores, int32_t* out_counts,
                                 cudaStream_t st) {
    rerank_select_kernel<<<Q, 256, 0, st>>>(scores, boosts, lens, L, top_k, min_score, out_idx, out_scores, out_counts);
    return cudaGetLastError();
}

int rrf_max_entries() { return kRrfMaxEntries; }

cudaError_t rrf_launch(const int32_t* ids, const double* weights, int Q, int R, int L, int rrf_k, int top,
                       int32_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st) {
    const int np = (R * L + 7) & ~7;
    const size_t smem = (size_t)np * (8 + 4 + 2 + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_dynamic_smem_of(rrf_kernel, (size_t)(smem));
        if (e != cudaSuccess) return e;
    }
    rrf_kernel<<<Q, 256, smem, st>>>(ids, weights, R, L, rrf_k, top, out_ids, out_scores, out_counts);
    return cudaGetLastError();
}

}  // namespace b200rag
