// dense_scan.cu — small-batch (B <= 4) query x corpus similarity with a fused
// running top-k, for the HBM-bound regime.
//
// Replaces the distance computation + selection inside collection.query(...)
// (reference call sites: src/rag/retriever.py:215-220, 380-385).
//
// One persistent CTA per SM.  Warp 0 is the producer: it streams contiguous
// row tiles (<= 32 KB) from HBM into a shared-memory ring with TMA 1-D bulk
// copies (cp.async.bulk + mbarrier complete_tx).  The consumer warps take rows
// of the landed tile: each lane owns a fixed 16-byte slice of every 512-byte
// chunk of the row (conflict-free LDS.128), keeps the matching slice of the
// NQ fp32 queries in registers, accumulates in fp32, and butterflies the lane
// partials.  The score never goes to HBM: it is packed with the row id into a
// u64 key and offered to the warp's running top-KP (WarpTopK).  Each warp
// finally writes its KP best keys; dense_select.cu merges and refines them.
//
// Algorithmic bytes: n * dim * sizeof(dtype) per launch, read exactly once.
#include "common.cuh"
#include "kernels.h"

namespace b200rag {

template <int DT>
struct Elem;
template <>
struct Elem<RAG_F32> { static constexpr int kBytes = 4; static constexpr int kPerLane = 4; };
template <>
struct Elem<RAG_BF16> { static constexpr int kBytes = 2; static constexpr int kPerLane = 8; };
template <>
struct Elem<RAG_F16> { static constexpr int kBytes = 2; static constexpr int kPerLane = 8; };

// unpack one 16-byte slice into fp32 values
template <int DT>
__device__ __forceinline__ void unpack16(const uint4& v, float* f) {
    if constexpr (DT == RAG_F32) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    } else if constexpr (DT == RAG_BF16) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    } else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = p.x; f[2 * i + 1] = p.y;
        }
    }
}

// MODE 0: running top-KP per warp.  MODE 1: collect every row whose filter
// score is >= tau[q] (fallback pass after a failed margin check).
template <int DT, int NCH, int NQ, int CW, int MODE>
__global__ void __launch_bounds__(32 + CW * 32, 1)
dense_scan_kernel(ScanParams p) {
    constexpr int EPL = Elem<DT>::kPerLane;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + p.n_stages;
    uint8_t* stage_base = smem + 128;
    uint64_t* cand_base = reinterpret_cast<uint64_t*>(stage_base + (size_t)p.n_stages * p.tile_bytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // device-driven fallback pass: the number of queries is only known on the device (block-uniform early exit)
    int n_queries = p.n_queries;
    if (p.n_active_dev != nullptr) {
        n_queries = *p.n_active_dev - p.active_first;
        if (n_queries > NQ) n_queries = NQ;
        if (n_queries <= 0) return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t n_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                int64_t row0 = t * p.tile_rows;
                int64_t rows = p.n_rows - row0 < p.tile_rows ? p.n_rows - row0 : p.tile_rows;
                uint32_t bytes = (uint32_t)(rows * p.row_bytes);
                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                bulk_g2s(stage_base + (size_t)stage * p.tile_bytes,
                         reinterpret_cast<const uint8_t*>(p.rows) + row0 * p.row_bytes, bytes, &full_bar[stage]);
                if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ===== consumers =====
    const int cw = warp - 1;
    // this lane's slice of each query, in registers
    float qreg[NQ][NCH * EPL];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                int col = c * (32 * EPL) + lane * EPL + e;
                qreg[qi][c * EPL + e] = (qi < n_queries && col < p.dim) ? p.q[(size_t)qi * p.dim + col] : 0.f;
            }
        }
    }

    WarpTopK top[NQ];
    if constexpr (MODE == 0) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi)
            top[qi].init(cand_base + ((size_t)cw * NQ + qi) * 2 * p.kp, p.kp, lane);
    }
    float tau[NQ];
    if constexpr (MODE == 1) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) tau[qi] = qi < n_queries ? p.tau[qi] : 3.0e38f;
    }

    int stage = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        const int64_t row0 = t * p.tile_rows;
        const int rows = (int)(p.n_rows - row0 < p.tile_rows ? p.n_rows - row0 : p.tile_rows);
        const uint8_t* tile = stage_base + (size_t)stage * p.tile_bytes;
        for (int r = cw; r < rows; r += CW) {
            const uint8_t* rowp = tile + (size_t)r * p.row_bytes;
            uint4 v[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                int off = c * 512 + lane * 16;
                v[c] = off < p.row_bytes ? *reinterpret_cast<const uint4*>(rowp + off) : make_uint4(0, 0, 0, 0);
            }
            float acc[NQ][4];
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) acc[qi][0] = acc[qi][1] = acc[qi][2] = acc[qi][3] = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float f[EPL];
                unpack16<DT>(v[c], f);
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
                    for (int e = 0; e < EPL; ++e)
                        acc[qi][e & 3] = fmaf(f[e], qreg[qi][c * EPL + e], acc[qi][e & 3]);
                }
            }
            const uint32_t grow = (uint32_t)(row0 + r);
            bool allowed = bitmap_test(p.allow, grow);
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                float s = (acc[qi][0] + acc[qi][1]) + (acc[qi][2] + acc[qi][3]);
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (qi < n_queries && allowed) {
                    if constexpr (MODE == 0) {
                        top[qi].push(make_key(s, grow), lane);
                    } else {
                        if (s >= tau[qi] && lane == 0) {
                            unsigned idx = atomicAdd(&p.collect_count[qi], 1u);
                            if (idx < (unsigned)p.collect_cap) p.collect_rows[(size_t)qi * p.collect_cap + idx] = grow;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
    }

    if constexpr (MODE == 0) {
        // per-CTA merge: every consumer warp finishes its list, then consumer warp 0 folds the other
        // warps' KP best keys into its own and publishes ONE list per (query, CTA)
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi)
            if (qi < n_queries) top[qi].finish(lane);
        asm volatile("bar.sync 1, %0;" ::"r"(CW * 32) : "memory");       // consumer warps only
        if (cw == 0) {
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                if (qi >= n_queries) break;
                for (int w = 1; w < CW; ++w) {
                    const uint64_t* other = cand_base + ((size_t)w * NQ + qi) * 2 * p.kp;
                    for (int i0 = 0; i0 < p.kp; i0 += kWarp) {
                        const int i = i0 + lane;
                        top[qi].offer(i < p.kp ? other[i] : 0ull, lane);
                    }
                }
                top[qi].finish(lane);
                uint64_t* out = p.cand + ((size_t)qi * p.n_lists + blockIdx.x) * p.kp;
                for (int i = lane; i < p.kp; i += kWarp) out[i] = i < top[qi].n ? top[qi].buf[i] : 0ull;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
template <int DT, int NCH, int NQ, int MODE>
static cudaError_t launch_inst(const ScanParams& p, int grid, size_t smem, cudaStream_t st) {
    constexpr int CW = NQ <= 2 ? 16 : 8;
    auto kern = dense_scan_kernel<DT, NCH, NQ, CW, MODE>;
    cudaError_t e = ensure_dynamic_smem_of(kern, (size_t)(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, 32 + CW * 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int DT, int NCH, int MODE>
static cudaError_t launch_nq(const ScanParams& p, int nq_t, int grid, size_t smem, cudaStream_t st) {
    switch (nq_t) {
        case 1: return launch_inst<DT, NCH, 1, MODE>(p, grid, smem, st);
        case 2: return launch_inst<DT, NCH, 2, MODE>(p, grid, smem, st);
        default: return launch_inst<DT, NCH, 4, MODE>(p, grid, smem, st);
    }
}

template <int DT, int MODE>
static cudaError_t launch_nch(const ScanParams& p, int nch, int nq_t, int grid, size_t smem, cudaStream_t st) {
    switch (nch) {
        case 1: return launch_nq<DT, 1, MODE>(p, nq_t, grid, smem, st);
        case 2: return launch_nq<DT, 2, MODE>(p, nq_t, grid, smem, st);
        case 4: return launch_nq<DT, 4, MODE>(p, nq_t, grid, smem, st);
        default: return launch_nq<DT, 8, MODE>(p, nq_t, grid, smem, st);
    }
}

static int scan_consumer_warps(int nq_t) { return nq_t <= 2 ? 16 : 8; }
int scan_nq_template(int n_queries) { return n_queries <= 1 ? 1 : (n_queries <= 2 ? 2 : 4); }

// Fills the derived fields of p (tile geometry, stages) and returns the dynamic
// shared-memory size; n_lists = grid (the CTA merges its warps' lists).
size_t scan_plan(ScanParams& p, int dtype, int sm_count, int smem_limit, int* grid_out, int* nch_out) {
    const int esz = dtype == RAG_F32 ? 4 : 2;
    p.row_bytes = p.dim * esz;
    int nch = (p.row_bytes + 511) / 512;
    nch = nch <= 1 ? 1 : (nch <= 2 ? 2 : (nch <= 4 ? 4 : 8));
    *nch_out = nch;
    int tile_rows = 32768 / p.row_bytes;
    if (tile_rows < 8) tile_rows = 8;
    p.tile_rows = tile_rows;
    p.tile_bytes = tile_rows * p.row_bytes;
    const int nq_t = p.nq_t;
    const int cw = scan_consumer_warps(nq_t);
    size_t cand_bytes = p.mode == 0 ? (size_t)cw * nq_t * 2 * p.kp * sizeof(uint64_t) : 0;
    int stages = (int)((smem_limit - 128 - (long)cand_bytes) / p.tile_bytes);
    if (stages > 6) stages = 6;
    if (stages < 2) return 0;
    p.n_stages = stages;
    int64_t n_tiles = (p.n_rows + tile_rows - 1) / tile_rows;
    int grid = (int)(n_tiles < sm_count ? (n_tiles > 0 ? n_tiles : 1) : sm_count);
    *grid_out = grid;
    p.n_lists = grid;            // one merged list per (query, CTA)
    return 128 + (size_t)stages * p.tile_bytes + cand_bytes;
}

cudaError_t scan_launch(const ScanParams& p, int dtype, int nch, int grid, size_t smem, cudaStream_t st) {
    const int nq_t = p.nq_t;
    if (p.mode == 0) {
        switch (dtype) {
            case RAG_F32: return launch_nch<RAG_F32, 0>(p, nch, nq_t, grid, smem, st);
            case RAG_BF16: return launch_nch<RAG_BF16, 0>(p, nch, nq_t, grid, smem, st);
            default: return launch_nch<RAG_F16, 0>(p, nch, nq_t, grid, smem, st);
        }
    } else {
        switch (dtype) {
            case RAG_F32: return launch_nch<RAG_F32, 1>(p, nch, nq_t, grid, smem, st);
            case RAG_BF16: return launch_nch<RAG_BF16, 1>(p, nch, nq_t, grid, smem, st);
            default: return launch_nch<RAG_F16, 1>(p, nch, nq_t, grid, smem, st);
        }
    }
}

}  // namespace b200rag
