// common.cuh — device helpers shared by the b200rag kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "b200rag kernels are written for sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif
#endif

namespace b200rag {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// candidate keys.  A dense candidate is one u64: high word = order-preserving
// image of the fp32 filter score, low word = ~row, so that "larger key" ==
// "(score desc, row asc) ranks earlier".  Key 0 is "empty".
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return ((uint64_t)f32_ordered(score) << 32) | (uint64_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~(uint32_t)key; }
__host__ __device__ __forceinline__ float key_score(uint64_t key) { return f32_from_ordered((uint32_t)(key >> 32)); }

// ---------------------------------------------------------------------------
// small PTX wrappers: mbarrier + TMA 1-D bulk copy
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// the same with the barrier given as a shared-space address (computed once outside a hot loop)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or ~1 ms passes, so a
// waiting warp — a producer in front of a full ring — does not burn issue slots polling)
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
// Bounded: a protocol bug traps (-> CUDA error) after a few seconds instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_a(bar, parity)) return;
    int spins = 0;
    while (!mbar_try_wait_a(bar, parity)) {
        if (++spins > (1 << 22)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// global -> shared bulk copy (TMA, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------
// warp-cooperative bitonic sort (descending) of n (power of two, >= 32) keys
// held in shared memory.  Ends with __syncwarp().
// ---------------------------------------------------------------------------
template <typename K>
__device__ __forceinline__ void warp_bitonic_desc(K* buf, int n, int lane) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < (n >> 1); t += kWarp) {
                int lo = 2 * t - (t & (stride - 1));      // index with bit `stride` clear
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                K a = buf[lo], b = buf[hi];
                bool swap = desc ? (a < b) : (b < a);
                if (swap) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncwarp();
}

// block-cooperative version (all threads of the CTA call it)
template <typename K>
__device__ __forceinline__ void block_bitonic_desc(K* buf, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                K a = buf[lo], b = buf[hi];
                bool swap = desc ? (a < b) : (b < a);
                if (swap) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// per-warp running top-KP over keys of type K (K{} == "empty", ordered by
// operator< / operator>): append survivors to a 2*KP buffer in shared memory,
// prune (sort, keep KP, raise threshold) when it fills.  All lanes call every
// method; push() takes a warp-uniform key, offer() one key per lane.
// ---------------------------------------------------------------------------
template <typename K>
struct WarpTopKT {
    K* buf;     // 2*KP entries in shared memory
    int kp;     // power of two >= 16
    int n;      // entries in buf (warp-uniform)
    K thr;      // keys <= thr cannot enter the top-KP

    __device__ __forceinline__ void init(K* b, int kp_, int lane) {
        buf = b; kp = kp_; n = 0; thr = K{};
        for (int i = lane; i < 2 * kp_; i += kWarp) b[i] = K{};
        __syncwarp();
    }
    __device__ __forceinline__ void prune(int lane) {
        // sort the smallest power of two that covers the n live entries (entries past it are never read
        // before they are overwritten by later appends or cleared by finish())
        int cap = kWarp < 2 * kp ? kWarp : 2 * kp;
        while (cap < n) cap <<= 1;
        for (int i = n + lane; i < cap; i += kWarp) buf[i] = K{};
        warp_bitonic_desc(buf, cap, lane);
        if (n > kp) n = kp;
        if (n == kp) thr = buf[kp - 1];
    }
    __device__ __forceinline__ void push(K key, int lane) {
        if (key > thr) {
            if (lane == 0) buf[n] = key;
            ++n;
            if (n == 2 * kp) prune(lane);
        }
    }
    __device__ __forceinline__ void offer(K key, int lane) {
        bool pass = key > thr;
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int cnt = __popc(m);
            const int room = 2 * kp - n;
            const int rank = __popc(m & ((1u << lane) - 1));
            if (cnt <= room) {
                if (pass) buf[n + rank] = key;
                n += cnt;
                if (n == 2 * kp) prune(lane);
                break;
            }
            if (pass && rank < room) { buf[n + rank] = key; pass = false; }
            n = 2 * kp;
            prune(lane);
            pass = pass && key > thr;
            m = __ballot_sync(0xffffffffu, pass);
        }
    }
    // final prune; afterwards buf[0..n) is sorted descending and buf[n..2kp) is empty
    __device__ __forceinline__ void finish(int lane) {
        prune(lane);
        for (int i = n + lane; i < 2 * kp; i += kWarp) buf[i] = K{};
        __syncwarp();
    }
};
using WarpTopK = WarpTopKT<uint64_t>;

__device__ __forceinline__ bool bitmap_test(const uint8_t* __restrict__ bm, uint32_t row) {
    return bm == nullptr || ((bm[row >> 3] >> (row & 7)) & 1);
}

}  // namespace b200rag
