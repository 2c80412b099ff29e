// corpus.cu — write path of the chunk-embedding matrix: dtype conversion on
// upload, widening on download, row compaction (delete), the running max row
// norm the filter bound needs, and a deterministic on-device synthetic fill.
//
// Reference write path this backs: collection.add(ids, documents, embeddings,
// metadatas) at src/processing/create_chromadb_index.py:374-379 and
// src/processing/ingest_enterprise.py:241-246; collection.delete at
// src/processing/ingest_enterprise.py:272,304.
#include "common.cuh"
#include <map>
#include <mutex>

#include "kernels.h"

namespace b200rag {

template <int DT>
__device__ __forceinline__ void store_elem(void* dst, size_t i, float v) {
    if constexpr (DT == RAG_F32) reinterpret_cast<float*>(dst)[i] = v;
    else if constexpr (DT == RAG_BF16) reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);
}
template <int DT>
__device__ __forceinline__ float load_elem(const void* src, size_t i) {
    if constexpr (DT == RAG_F32) return reinterpret_cast<const float*>(src)[i];
    else if constexpr (DT == RAG_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
    else return __half2float(reinterpret_cast<const __half*>(src)[i]);
}

template <int DT>
__global__ void convert_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) store_elem<DT>(dst, i, src[i]);
}
template <int DT>
__global__ void widen_kernel(const void* __restrict__ src, float* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = load_elem<DT>(src, i);
}

static int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t convert_rows_launch(const float* src, void* dst, int dtype, int64_t n, cudaStream_t st) {
    int g = grid_for(n, 256);
    if (dtype == RAG_F32) convert_kernel<RAG_F32><<<g, 256, 0, st>>>(src, dst, n);
    else if (dtype == RAG_BF16) convert_kernel<RAG_BF16><<<g, 256, 0, st>>>(src, dst, n);
    else convert_kernel<RAG_F16><<<g, 256, 0, st>>>(src, dst, n);
    return cudaGetLastError();
}
cudaError_t widen_rows_launch(const void* src, int dtype, float* dst, int64_t n, cudaStream_t st) {
    int g = grid_for(n, 256);
    if (dtype == RAG_F32) widen_kernel<RAG_F32><<<g, 256, 0, st>>>(src, dst, n);
    else if (dtype == RAG_BF16) widen_kernel<RAG_BF16><<<g, 256, 0, st>>>(src, dst, n);
    else widen_kernel<RAG_F16><<<g, 256, 0, st>>>(src, dst, n);
    return cudaGetLastError();
}

// max over rows of ||x||_2 (an upper bound: inflated by 1e-5 to cover the fp32 sum)
template <int DT>
__global__ void row_norm_max_kernel(const void* __restrict__ rows, int64_t n_rows, int dim, float* max_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        float s = 0.f;
        for (int c = lane; c < dim; c += 32) {
            float v = load_elem<DT>(rows, (size_t)r * dim + c);
            s = fmaf(v, v, s);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        best = fmaxf(best, s);
    }
    if (lane == 0 && best > 0.f) {
        float nrm = sqrtf(best) * 1.00001f;
        atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(nrm));   // positive floats order as ints
    }
}

cudaError_t row_norm_max_launch(const void* rows, int dtype, int64_t n_rows, int dim, float* max_norm,
                                cudaStream_t st) {
    int g = grid_for(n_rows * 32, 256);
    if (dtype == RAG_F32) row_norm_max_kernel<RAG_F32><<<g, 256, 0, st>>>(rows, n_rows, dim, max_norm);
    else if (dtype == RAG_BF16) row_norm_max_kernel<RAG_BF16><<<g, 256, 0, st>>>(rows, n_rows, dim, max_norm);
    else row_norm_max_kernel<RAG_F16><<<g, 256, 0, st>>>(rows, n_rows, dim, max_norm);
    return cudaGetLastError();
}

// Deterministic synthetic rows.  value(seed,row,col) = splitmix64 finaliser of
// seed + row*G1 + col*G2 -> top 24 bits -> uniform in [-1,1); the row is then
// L2-normalised (fp64 sum of squares in the canonical lane order, fp64 divide)
// and rounded to the storage dtype.  b200rag/synth.py is the numpy twin.
__host__ __device__ __forceinline__ float synth_value(uint64_t seed, uint64_t row, uint64_t col) {
    uint64_t z = seed + row * 0x9E3779B97F4A7C15ULL + col * 0xD1B54A32D192ED03ULL;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (float)(uint32_t)(z >> 40) * (1.0f / 8388608.0f) - 1.0f;
}

template <int DT>
__global__ void fill_synthetic_kernel(void* __restrict__ rows, int64_t row0, int64_t n_rows, int dim, uint64_t seed,
                                      int64_t gen_row0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const uint64_t grow = (uint64_t)(gen_row0 + r);
        double p = 0.0;
        for (int j = 0; j < dim / 32; ++j) {
            double v = (double)synth_value(seed, grow, (uint64_t)(32 * j + lane));
            p = __fma_rn(v, v, p);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_down_sync(0xffffffffu, p, off));
        p = __shfl_sync(0xffffffffu, p, 0);
        const double nrm = __dsqrt_rn(p);
        for (int j = 0; j < dim / 32; ++j) {
            const int c = 32 * j + lane;
            double v = (double)synth_value(seed, grow, (uint64_t)c);
            store_elem<DT>(rows, (size_t)(row0 + r) * dim + c, __double2float_rn(__ddiv_rn(v, nrm)));
        }
    }
}

cudaError_t fill_synthetic_launch(void* rows, int dtype, int64_t row0, int64_t n_rows, int dim, uint64_t seed,
                                  int64_t gen_row0, cudaStream_t st) {
    int g = grid_for(n_rows * 32, 256);
    if (dtype == RAG_F32) fill_synthetic_kernel<RAG_F32><<<g, 256, 0, st>>>(rows, row0, n_rows, dim, seed, gen_row0);
    else if (dtype == RAG_BF16) fill_synthetic_kernel<RAG_BF16><<<g, 256, 0, st>>>(rows, row0, n_rows, dim, seed, gen_row0);
    else fill_synthetic_kernel<RAG_F16><<<g, 256, 0, st>>>(rows, row0, n_rows, dim, seed, gen_row0);
    return cudaGetLastError();
}

__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                   const int64_t* __restrict__ keep, int64_t nkeep, int row_bytes) {
    for (int64_t i = blockIdx.x; i < nkeep; i += gridDim.x) {
        const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)keep[i] * row_bytes);
        uint4* d = reinterpret_cast<uint4*>(dst + (size_t)i * row_bytes);
        for (int c = threadIdx.x; c < row_bytes / 16; c += blockDim.x) d[c] = s[c];
    }
}

cudaError_t gather_rows_launch(const void* src, void* dst, const int64_t* keep, int64_t nkeep, int row_bytes,
                               cudaStream_t st) {
    int g = (int)(nkeep < 148 * 8 ? (nkeep > 0 ? nkeep : 1) : 148 * 8);
    gather_rows_kernel<<<g, 128, 0, st>>>(reinterpret_cast<const uint8_t*>(src), reinterpret_cast<uint8_t*>(dst), keep,
                                          nkeep, row_bytes);
    return cudaGetLastError();
}


// ---- per-(device, kernel) cache of the dynamic shared-memory opt-in ----------------------------------------------
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
    struct Key {
        const void* fn;
        int dev;
        bool operator<(const Key& o) const { return fn < o.fn || (fn == o.fn && dev < o.dev); }
    };
    static std::mutex mu;
    static std::map<Key, size_t> seen;            // the largest size already set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    auto it = seen.find(Key{kernel, dev});
    if (it != seen.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) seen[Key{kernel, dev}] = bytes;
    return e;
}

}  // namespace b200rag
