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

int rrf_max_entries() { return kRrfMaxEntries; }

cudaError_t rrf_launch(const int32_t* ids, const double* weights, int Q, int R, int L, int rrf_k, int top,
                       int32_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st) {
    const int np = (R * L + 7) & ~7;
    const size_t smem = (size_t)np * (8 + 4 + 2 + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    rrf_kernel<<<Q, 256, smem, st>>>(ids, weights, R, L, rrf_k, top, out_ids, out_scores, out_counts);
    return cudaGetLastError();
}

}  // namespace b200rag
