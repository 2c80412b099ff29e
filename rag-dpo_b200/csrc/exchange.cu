// exchange.cu — the multi-GPU exchange step over NVLink peer memory, without a collective library call.
//
// One process per GPU: every rank holds the exact local top-k of the same query batch and the
// global top-k is the merge of the G lists.  Instead of an NCCL all-gather followed by a merge kernel, each rank
// STORES its (score, global id) block straight into every peer's gather buffer (cudaIpc-mapped device memory,
// NVLink 5 / NVSwitch P2P), raises an epoch flag there, and the merge kernel of every rank waits for the G
// flags of the epoch before it merges.  Two launches on the caller's stream, no host synchronisation, no
// rendezvous in a communication library.
//
// Buffer of a rank (allocated locally, mapped by every peer):
//   payload[2][G][slot_bytes]   parity (epoch & 1) x source rank x (scores f64[B*k] | ids i64[B*k])
//   flags  [2][G]               u64: last epoch whose payload from that source rank is complete
// Reuse is safe with two parities: a rank can raise its flag for epoch e+1 only after its own merge of epoch e
// (stream order), i.e. after every rank's flag for e — so nobody is still reading parity (e+2)&1 == e&1 when a
// fast rank overwrites it for e+2, because that fast rank first had to see everyone's flag for e+1.
#include "common.cuh"
#include "kernels.h"

namespace b200rag {

__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid = (blocks_per_peer, G): copy my payload into slot `rank` of peer blockIdx.y; the last block to finish
// (device-wide counter) publishes the epoch flag in every peer
__global__ void __launch_bounds__(256)
exchange_push_kernel(ExchangeDev ex, const double* __restrict__ my_scores, const int64_t* __restrict__ my_ids,
                     const int32_t* __restrict__ my_rows, int64_t row_lo, const int32_t* __restrict__ my_counts, int k,
                     int64_t n_elems /* B*k */, uint64_t epoch, int vec_ok) {
    const int peer = blockIdx.y;
    const int parity = (int)(epoch & 1);
    uint8_t* dst = ex.peer_base[peer] + ((size_t)parity * ex.world + ex.rank) * ex.slot_bytes;
    // n_elems doubles then n_elems int64: both 8-byte items; 16-byte vectors when the count is even
    uint64_t* d64 = reinterpret_cast<uint64_t*>(dst);
    const uint64_t* s0 = reinterpret_cast<const uint64_t*>(my_scores);
    const uint64_t* s1 = reinterpret_cast<const uint64_t*>(my_ids);
    const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (vec_ok) {           // even element count, 16-byte aligned scores, 8-byte aligned rows (checked by the launcher)
        // pairs: 16-byte loads of the scores / 8-byte loads of two local rows, 16-byte stores into the peer's buffer
        const int64_t half = n_elems >> 1;
        ulonglong2* d128 = reinterpret_cast<ulonglong2*>(dst);
        const ulonglong2* sc2 = reinterpret_cast<const ulonglong2*>(my_scores);
        const int2* rw2 = reinterpret_cast<const int2*>(my_rows);
#pragma unroll 4
        for (int64_t i = tid0; i < n_elems; i += nthr) {
            ulonglong2 v;
            if (i < half) {
                v = sc2[i];
            } else {
                const int64_t pr = i - half, e = 2 * pr;
                const int2 r = rw2[pr];
                v.x = (unsigned long long)(r.x < 0 ? (int64_t)-1 : (int64_t)r.x + row_lo);
                v.y = (unsigned long long)(r.y < 0 ? (int64_t)-1 : (int64_t)r.y + row_lo);
                if (my_counts != nullptr) {      // (see below: an unfinished query is announced with id -2 in its first slot)
                    if (e % k == 0 && my_counts[e / k] < 0) v.x = (unsigned long long)(int64_t)-2;
                    if ((e + 1) % k == 0 && my_counts[(e + 1) / k] < 0) v.y = (unsigned long long)(int64_t)-2;
                }
            }
            d128[i] = v;
        }
    } else
    for (int64_t i = tid0; i < 2 * n_elems; i += nthr) {
        uint64_t v;
        if (i < n_elems) {
            v = s0[i];
        } else if (my_rows) {               // local int32 rows -> global ids on the way out (-1 stays padding)
            const int64_t e = i - n_elems;
            const int32_t r = my_rows[e];
            v = (uint64_t)(r < 0 ? (int64_t)-1 : (int64_t)r + row_lo);
            // a query this rank could not finish exactly (count < 0: more deep ties than the stream-ordered fallback
            // serves) is announced to every rank with id -2 in its first slot: all ranks mark it and redo it together
            if (my_counts != nullptr && e % k == 0 && my_counts[e / k] < 0) v = (uint64_t)(int64_t)-2;
        } else {
            v = s1[i - n_elems];
        }
        d64[i] = v;
    }
    __syncthreads();                                // the block's stores happen-before thread 0's fence
    if (threadIdx.x == 0) {
        __threadfence_system();                     // ... which orders them before the counter / the flags (cumulative)
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned prev = atomicAdd(ex.done_counter, 1u);
        if (prev == total - 1) {
            *ex.done_counter = 0u;                  // ready for the next launch (stream-ordered)
            __threadfence_system();                 // every block's payload is ordered before the flags below
            // plain system-scope stores, issued back to back: one NVLink trip for all peers instead of one
            // release (= wait for the previous store to land) per peer
            for (int p = 0; p < ex.world; ++p)
                st_relaxed_sys_u64(ex.peer_flags[p] + (size_t)parity * ex.world + ex.rank, epoch);
        }
    }
}

// bounded wait for the G flags of this epoch in MY buffer (thread 0), then the caller merges.  A peer that does
// not arrive within the timeout does NOT take the CUDA context down: the block reports it (returns false, the
// kernel marks its query with count = -2 and raises the exchange's error word), the host sees it at its next
// synchronisation (rag_exchange_status).
__device__ __forceinline__ bool exchange_wait(const uint64_t* flags, int world, uint64_t epoch, long long timeout_cycles,
                                              unsigned* err_word) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        for (int r = 0; r < world && ok; ++r) {
            if (ld_acquire_sys_u64(flags + r) >= epoch) continue;
            const long long t0 = clock64();
            while (ld_acquire_sys_u64(flags + r) < epoch) {
                if (clock64() - t0 > timeout_cycles) { ok = 0; break; }
                __nanosleep(200);
            }
        }
        if (!ok) atomicExch(err_word, 1u);
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

struct XKey {
    double s;
    int64_t id;
    __device__ __forceinline__ bool operator<(const XKey& o) const { return s < o.s || (s == o.s && id > o.id); }
};

// one CTA per query: wait for the epoch, then G sorted (score, id) lists -> global top-k (ties -> lowest id)
__global__ void __launch_bounds__(256)
exchange_merge_kernel(ExchangeDev ex, int B, int k, uint64_t epoch, int nsort, double* __restrict__ out_scores,
                      int64_t* __restrict__ out_ids, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    XKey* ek = reinterpret_cast<XKey*>(sm_raw);
    __shared__ int s_count;
    const int parity = (int)(epoch & 1);
    const int b = blockIdx.x;
    if (threadIdx.x == 0) s_count = 0;
    if (!exchange_wait(ex.my_flags + (size_t)parity * ex.world, ex.world, epoch, ex.timeout_cycles, ex.done_counter + 1)) {
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            out_ids[(size_t)b * k + i] = -1;
            out_scores[(size_t)b * k + i] = 0.0;
        }
        if (threadIdx.x == 0) out_counts[b] = -2;
        return;
    }
    const uint8_t* base = ex.my_base + (size_t)parity * ex.world * ex.slot_bytes;
    const int64_t n_elems = (int64_t)B * k;
    const int n_e = ex.world * k;
    int local = 0;
    for (int i = threadIdx.x; i < n_e; i += blockDim.x) {
        XKey e;
        e.s = -INFINITY; e.id = INT64_MAX;
        const int g = i / k, j = i % k;
        const double* sc = reinterpret_cast<const double*>(base + (size_t)g * ex.slot_bytes);
        const int64_t* ids = reinterpret_cast<const int64_t*>(sc + n_elems);
        const int64_t id = ids[(size_t)b * k + j];
        if (id >= 0) { e.s = sc[(size_t)b * k + j]; e.id = id; ++local; }
        else if (id == -2) local += 1 << 20;          // some rank could not finish this query exactly
        ek[i] = e;
    }
    atomicAdd(&s_count, local);
    __syncthreads();
    const bool unresolved = s_count >= (1 << 20);
    const int n_valid = s_count & ((1 << 20) - 1);
    const int nout = n_valid < k ? n_valid : k;
    // Every list arrives sorted (score desc, id asc) and the ids of different ranks are disjoint: the global rank of
    // entry j of list g is j + the number of entries of every other list that precede it — one binary search per
    // other list, no sort, no further barrier.
    for (int i = threadIdx.x; i < n_e; i += blockDim.x) {
        const XKey e = ek[i];
        if (e.id == INT64_MAX) continue;
        const int g = i / k;
        int rank = i - g * k;
        for (int o = 0; o < ex.world && rank < k; ++o) {
            if (o == g) continue;
            const XKey* L = ek + o * k;
            int lo = 0, hi = k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (e < L[mid]) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_ids[(size_t)b * k + rank] = e.id;
            out_scores[(size_t)b * k + rank] = e.s;
        }
    }
    for (int i = nout + threadIdx.x; i < k; i += blockDim.x) {
        out_ids[(size_t)b * k + i] = -1;
        out_scores[(size_t)b * k + i] = 0.0;
    }
    if (threadIdx.x == 0) out_counts[b] = unresolved ? -1 : nout;
}

cudaError_t exchange_push_launch(const ExchangeDev& ex, const double* my_scores, const int64_t* my_ids,
                                 const int32_t* my_rows, int64_t row_lo, const int32_t* my_counts, int B, int k,
                                 uint64_t epoch, cudaStream_t st) {
    const int64_t n_elems = (int64_t)B * k;
    int bx = (int)((2 * n_elems + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dim3 grid(bx, ex.world);
    const int vec_ok = my_rows != nullptr && (n_elems & 1) == 0 && (reinterpret_cast<uintptr_t>(my_scores) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(my_rows) & 7) == 0;
    exchange_push_kernel<<<grid, 256, 0, st>>>(ex, my_scores, my_ids, my_rows, row_lo, my_counts, k, n_elems, epoch, vec_ok);
    return cudaGetLastError();
}

cudaError_t exchange_merge_launch(const ExchangeDev& ex, int B, int k, uint64_t epoch, double* out_scores,
                                  int64_t* out_ids, int32_t* out_counts, cudaStream_t st) {
    const int nsort = ex.world * k;                  // (entries held in shared memory; they are ranked, not sorted)
    const size_t smem = (size_t)nsort * sizeof(XKey);
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_dynamic_smem_of(exchange_merge_kernel, (size_t)(smem));
        if (e != cudaSuccess) return e;
    }
    exchange_merge_kernel<<<B, 256, smem, st>>>(ex, B, k, epoch, nsort, out_scores, out_ids, out_counts);
    return cudaGetLastError();
}


}  // namespace b200rag
