// dense_gemm.cu — batched query x corpus similarity as a tcgen05 tensor-core
// contraction with a fused running top-k (the score matrix never exists in HBM).
//
// Replaces the distance computation + selection inside collection.query(...)
// (reference call sites: src/rag/retriever.py:215-220, 380-385) for batches of
// queries (the reference issues them one at a time).
//
// Shape of the contraction (per CTA tile):
//   D[128 queries x 256 rows] (fp32, TMEM) += Q[128 x 64] (bf16, smem) * X[256 x 64]^T (bf16, smem)
// Queries are the M side so that one TMEM lane == one query: each epilogue
// thread owns one query and keeps its threshold / candidate count in registers.
//
// Warp roles (192 threads, 1 CTA per SM, persistent over row tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor (128B-swizzled boxes) into a 4-stage ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128,N=256,K=16) x 4 per stage;
//               tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld the accumulator (double-buffered: 2 x 256 TMEM columns, so the
//               select of tile i overlaps the MMAs of tile i+1) and run the fused select.
//
// Fused select, two launches of the same kernel:
//   SAMPLE pass  over a fraction 8/E of the rows, E = max(1024, 4*kp) (column-granular): every thread keeps
//                the 8 best scores of its query in registers (inserted branch-free from the LDTM.x32
//                registers); merged per query (sample_tau_kernel), the 8th best sample score tau_q is a
//                VALID lower bound of the corpus-wide 8th best score.  A sample of <= 64 rows per CTA stays
//                RESIDENT in shared memory for all query blocks and is multiplied with a 64-wide MMA.
//   MAIN pass    over all tiles (row tile outer, query block inner: a corpus tile is fetched from HBM
//                once and re-read from L2 by the other query blocks): a branch-free compare mask per
//                32-column chunk against tau_q; the ~E survivors per query are parked per thread and
//                flushed once per tile (one atomic slot reservation) into that query's candidate list.
//                Everything with filter score >= tau_q is captured, so the candidate set provably
//                contains the top-k unless the list overflows (flagged -> exact fallback pass).
//
// Pair mode (NCTA = 2, main pass with an even number of query blocks): two CTAs of a cluster — the two SMs of
// one TPC — work on ONE 256-query x 256-row tile with tcgen05.mma.cta_group::2.  Each CTA stages its own 128
// queries and HALF of the row tile (16 KB + 16 KB per stage instead of 16 + 32), the leader CTA issues the
// MMAs, each CTA's TMEM receives its own 128 x 256 accumulator and runs the same fused select.  Operand
// traffic from L2 and shared-memory reads per SM drop by a third.
//
// Both operands are 16-bit: bf16 (a bf16 corpus, or the bf16 shadow of an fp32 corpus) or fp16 (an fp16 corpus,
// used as stored); queries are rounded to the same format by query_prep_kernel, which also returns the exact
// norm of the rounding residual.
// The result is only a FILTER: dense_select.cu re-scores the survivors with the canonical fp64 dot
// product over the stored values and checks the margin against the rigorous filter error bound.
//
// Algorithmic work: 2 * B * n * dim flops per call; HBM bytes: ONE pass over the bf16 rows (+ the sample).
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace b200rag {

constexpr int GT_M = 128;
constexpr int GT_N = 256;
constexpr int GT_K = 64;
constexpr int GT_A_BYTES = GT_M * GT_K * 2;   // 16 KB
constexpr int GT_B_BYTES = GT_N * GT_K * 2;   // 32 KB
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;
constexpr int GT_THREADS = 192;
constexpr int GT_EPI_WARPS = 4;
constexpr int kPend = 4;           // survivors a thread parks per tile before it must flush
constexpr int kMaxQBlocks = 32;   // per launch: 4096 queries (their tau / count live in shared memory)

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}

// ---- cluster (CTA pair) helpers ----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a location in this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id) issues it: the TMEM reads
// are ordered by tcgen05.fence::before_thread_sync, no cluster-scope release of this thread's global stores is needed
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a pair: data lands in THIS CTA's shared memory, the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                 uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
// the MMAs issued so far by this thread are complete -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 = 64 |
//       [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
    uint64_t d = (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32 at [4,6), A format at [7,10), B format at [10,13) (0 = f16, 1 = bf16),
// both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_16bit(int m, int n, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- the kernel ---------------------------------------------------------------------------------
constexpr int kSampleM = 8;       // order statistic taken from the sample pass

// Work items of one worker (a CTA, or a CTA pair in pair mode), identical in all three warp roles.
//   sample pass (MODE 0): query block outer, the CTA's sample tiles inner (the running top-m of a
//                         query lives in registers across tiles)
//   main pass   (MODE 1): row tile outer, query item inner: a corpus tile is fetched from HBM once
//                         and re-read from L2 for every query item (a query block; a PAIR of blocks in pair mode)
// tiles of this worker: t = worker + i * n_workers, i = 0, step, 2*step, ... (at most max_tiles)
template <int MODE, int NCTA>
struct WorkIter {
    int64_t n_tiles, i, i_first;
    int step, max_tiles, n_qitems;
    int worker, n_workers;
    int taken, qb;
    bool started;
    // MODE 1 tail: the n_tiles % n_workers leftover tiles are split by (tile, query item) ITEMS over all
    // workers, so that they finish within a few items of each other instead of a whole tile apart
    int64_t full_rounds, tail_item, tail_end;
    bool in_tail;
    __device__ __forceinline__ void init(const GemmParams& p) {
        n_tiles = (p.n_rows + GT_N - 1) / GT_N;
        step = MODE == 0 ? p.sample_step : 1;
        max_tiles = MODE == 0 ? p.sample_tiles : 0x7fffffff;
        n_qitems = (p.n_qblocks + NCTA - 1) / NCTA;
        worker = blockIdx.x / NCTA;
        n_workers = gridDim.x / NCTA;
        // the sample of CTA c starts at a pseudo-random one of its tiles, so that the union of the CTAs'
        // samples is spread over the whole corpus instead of being its first rows
        i_first = 0;
        if (MODE == 0) {
            const int64_t span = n_tiles / n_workers - (int64_t)(max_tiles - 1) * step;   // tiles every CTA has
            if (span > 1) i_first = (int64_t)((worker * 2654435761u) % (uint32_t)span);
        }
        i = i_first; taken = 0; qb = 0; started = false;
        in_tail = false;
        full_rounds = n_tiles / n_workers;
        const int64_t tail_items = (n_tiles - full_rounds * n_workers) * n_qitems;
        tail_item = tail_items * worker / n_workers;
        tail_end = tail_items * (worker + 1) / n_workers;
        if (MODE == 1 && p.balance_tail) max_tiles = (int)full_rounds; else tail_end = tail_item;
    }
    __device__ __forceinline__ bool tile_ok() const {
        return worker + i * n_workers < n_tiles && taken < max_tiles;
    }
    // advances to the next (tile, query item)
    __device__ __forceinline__ bool next(int64_t& t, int& qitem) {
        if (!in_tail) {
            if (!started) {
                started = true;
            } else if (MODE == 0) {
                i += step; ++taken;
                if (!tile_ok()) { i = i_first; taken = 0; ++qb; }
            } else {
                if (++qb == n_qitems) { qb = 0; i += step; ++taken; }
            }
            if (qb < n_qitems && tile_ok()) {
                t = worker + i * n_workers;
                qitem = qb;
                return true;
            }
            if (MODE == 0) return false;
            in_tail = true;
        } else {
            ++tail_item;
        }
        if (tail_item >= tail_end) return false;
        t = full_rounds * n_workers + tail_item / n_qitems;
        qitem = (int)(tail_item % n_qitems);
        return true;
    }
};

template <int MODE, int NCTA>     // MODE 0 = sample pass, 1 = main pass; NCTA 2 = CTA-pair MMA (main pass only)
__global__ void __launch_bounds__(GT_THREADS, 1)
dense_gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                       GemmParams p) {
    constexpr int B_ROWS = GT_N / NCTA;                   // rows of the tile this CTA stages
    constexpr int B_BYTES = B_ROWS * GT_K * 2;
    constexpr int STAGE_BYTES = GT_A_BYTES + B_BYTES;     // 48 KB, pair mode 32 KB
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the 128B-swizzled tiles need 1024-byte alignment in the shared address space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stages = smem;                                                    // n_stages * STAGE_BYTES
    // resident sample (MODE 0, small samples): the CTA's <= 64 sample rows stay in shared memory for all query
    // blocks (dim/64 blocks of 8 KB) and only the queries stream through a ring of 16 KB stages
    const bool resident = MODE == 0 && p.sample_resident;
    constexpr int RES_B_BYTES = 64 * GT_K * 2;                                 // 64 rows x 64 columns
    uint8_t* a_ring = stages + (size_t)(p.dim / GT_K) * RES_B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.ring_bytes);
    uint64_t* full_bar = bars;                     // [n_stages]   (pair mode: the leader's are used)
    uint64_t* empty_bar = bars + 8;                // [n_stages]
    uint64_t* tfull_bar = bars + 16;               // [2]
    uint64_t* tempty_bar = bars + 18;              // [2]          (pair mode: the leader's are used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    uint64_t* bres_bar = bars + 22;                // resident sample rows landed

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = p.dim / GT_K;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;      // 0 = leader of the pair
    WorkIter<MODE, NCTA> work;
    work.init(p);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], GT_EPI_WARPS * NCTA); }
        mbar_init(bres_bar, 1);
        mbar_fence_init();
    }
    if (NCTA == 2) cluster_sync_all();             // the peer's barriers exist before anything can signal them
    if (warp == 1) {   // TMEM: all 512 columns (two 128x256 fp32 accumulators); pair mode: in both CTAs at once
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int64_t t;
            int qi;
            if (resident) {
                WorkIter<MODE, NCTA> peek = work;
                if (peek.next(t, qi)) {            // the one sample tile of this CTA: its first 64 rows, all of K
                    mbar_arrive_expect_tx(bres_bar, num_kb * RES_B_BYTES);
                    for (int kb = 0; kb < num_kb; ++kb)
                        tma_load_2d(stages + (size_t)kb * RES_B_BYTES, &map_x, kb * GT_K, (int)(t * GT_N), bres_bar);
                }
            }
            while (work.next(t, qi)) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (resident) {
                        uint8_t* sa = a_ring + (size_t)stage * GT_A_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], GT_A_BYTES);
                        tma_load_2d(sa, &map_q, kb * GT_K, qi * GT_M, &full_bar[stage]);
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    uint8_t* sa = stages + (size_t)stage * STAGE_BYTES;
                    if (NCTA == 2) {
                        // both CTAs' bytes are counted on the leader's barrier, which only the leader arms
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        const uint32_t lbar = mapa_u32(&full_bar[stage], 0);
                        tma_load_2d_pair(sa, &map_q, kb * GT_K, (qi * 2 + (int)rank) * GT_M, lbar);
                        tma_load_2d_pair(sa + GT_A_BYTES, &map_x, kb * GT_K, (int)(t * GT_N) + (int)rank * B_ROWS, lbar);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(sa, &map_q, kb * GT_K, qi * GT_M, &full_bar[stage]);
                        tma_load_2d(sa + GT_A_BYTES, &map_x, kb * GT_K, (int)(t * GT_N), &full_bar[stage]);
                    }
                    if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair mode: the leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc_16bit(GT_M * NCTA, resident ? 64 : GT_N, p.fp16_operands ? 0u : 1u);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            int64_t t;
            int qi;
            bool rows_ready = !resident;
            for (; work.next(t, qi); ++it) {
                if (!rows_ready) { mbar_wait(bres_bar, 0); tc_fence_after(); rows_ready = true; }
                const uint32_t buf = it & 1;
                mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);      // epilogue(s) drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * GT_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);               // TMA bytes landed (in both CTAs)
                    tc_fence_after();
                    const uint8_t* sa = resident ? a_ring + (size_t)stage * GT_A_BYTES : stages + (size_t)stage * STAGE_BYTES;
                    const uint64_t a_desc = umma_desc_sw128(sa);
                    const uint64_t b_desc = umma_desc_sw128(resident ? stages + (size_t)kb * RES_B_BYTES : sa + GT_A_BYTES);
#pragma unroll
                    for (int k = 0; k < GT_K / 16; ++k) {             // +32 B per K=16 step inside the swizzle atom
                        if (NCTA == 2) tc_mma_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        else tc_mma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    // stage reusable (in both CTAs) once these MMAs retire
                    if (NCTA == 2) tc_commit_pair(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
                    if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                }
                if (NCTA == 2) tc_commit_pair(&tfull_bar[buf]); else tc_commit(&tfull_bar[buf]);   // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: fused select =====================
        // Fast path per 32-column chunk: one LDTM.x32 and a branch-free compare mask.  Survivors are rare,
        // so the slow path is warp-uniform: for every column that passed in ANY lane, re-read that single
        // column from TMEM (LDTM.x1) and let the lanes that own a survivor handle it.
        const int lg = warp & 3;                 // TMEM lane group this warp may read
        const int my = lg * 32 + lane;           // my query inside the query block == my TMEM lane
        const int n_chunks = MODE == 0 ? p.sample_chunks : GT_N / 32;
        const int n_tau_blocks = (p.n_qblocks + NCTA - 1) / NCTA * NCTA;
        float* s_tau = reinterpret_cast<float*>(bars + 24);            // [n_tau_blocks][128]
        // survivors of the current tile are parked here (per thread) and flushed with ONE atomic slot
        // reservation per thread and tile, so a warp waits for the atomic round trip once per tile
        uint64_t* s_pend = reinterpret_cast<uint64_t*>(s_tau + n_tau_blocks * GT_M) + (size_t)((warp - 2) * 32 + lane) * kPend;
        const uint32_t tempty_leader0 = NCTA == 2 ? mapa_u32(&tempty_bar[0], 0) : 0u;
        const uint32_t tempty_leader1 = NCTA == 2 ? mapa_u32(&tempty_bar[1], 0) : 0u;
        if (MODE == 1) {
            for (int b = (int)rank; b < n_tau_blocks; b += NCTA) {      // the query blocks this CTA serves
                const int q = b * GT_M + my;
                float tau = INFINITY;                                   // padded queries never pass
                if (q < p.n_queries) {
                    tau = -INFINITY;
                    if (p.tau_keys) {
                        const uint64_t tk = p.tau_keys[(size_t)q * kSampleM + (kSampleM - 1)];
                        if (tk != 0ull) tau = key_score(tk);
                    }
                }
                s_tau[b * GT_M + my] = tau;
            }
        }
        uint32_t it = 0;
        int cur_qb = -1;
        float best[kSampleM];                    // MODE 0: the best scores of my query, descending
        int64_t t;
        int qi;
        auto flush_sample = [&](int qblock) {
            const int q = qblock * GT_M + my;
            if (q >= p.n_queries) return;
            // keys only need to order by score here: a synthetic distinct id keeps them non-zero
            uint64_t* out = p.sample_keys + ((size_t)q * p.n_lists + blockIdx.x) * kSampleM;
#pragma unroll
            for (int i = 0; i < kSampleM; ++i)
                out[i] = best[i] == -INFINITY ? 0ull : make_key(best[i], (uint32_t)(blockIdx.x * kSampleM + i));
        };
        for (; work.next(t, qi); ++it) {
            const int qb = qi * NCTA + (int)rank;                // the query block of THIS CTA
            const int q = qb * GT_M + my;
            float tau = 0.f;
            uint64_t* my_list = nullptr;
            int* my_cnt = nullptr;
            int npend = 0;
            if (MODE == 0) {
                if (qb != cur_qb) {
                    if (cur_qb >= 0) flush_sample(cur_qb);
                    cur_qb = qb;
#pragma unroll
                    for (int i = 0; i < kSampleM; ++i) best[i] = -INFINITY;
                }
            } else {
                tau = s_tau[qb * GT_M + my];
                my_list = p.cand + (size_t)q * p.list_cap;      // ONE list per query, shared by all CTAs
                my_cnt = p.cand_cnt + q;
            }
            const uint32_t buf = it & 1;
            mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t row0 = (uint32_t)(t * GT_N);
            const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * GT_N;
            for (int c = 0; c < n_chunks; ++c) {
                uint32_t v[32];
                tc_ld32(tbase + c * 32, v);
                tc_wait_ld();
                const float thr = MODE == 0 ? (q < p.n_queries ? best[kSampleM - 1] : INFINITY) : tau;
                uint32_t m = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = __uint_as_float(v[j]);
                    const bool pass = MODE == 0 ? (s > thr) : (s >= thr);
                    m |= pass ? (1u << j) : 0u;
                }
                if constexpr (MODE == 0) {
                    // sample pass: early on every column beats the (still empty) list of every query, so the
                    // values are inserted straight from the registers of the LDTM.x32, branch-free: a column that
                    // does not count (outside the sample, past the last row, filtered out, or below my current
                    // 8th best) is replaced by -inf, which falls through the bubble
                    uint32_t valid = c == n_chunks - 1 ? p.sample_last_mask : 0xffffffffu;   // column-granular sample
                    const int64_t left = p.n_rows - (int64_t)row0 - c * 32;                   // rows from this chunk on
                    if (left < 32) valid &= left > 0 ? (1u << left) - 1u : 0u;
                    if (p.allow) {                                  // the chunk's 32 rows are 4 bytes of the bitmap
                        const int64_t byte0 = ((int64_t)row0 + c * 32) >> 3, n_bytes = (p.n_rows + 7) >> 3;
                        uint32_t w = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (byte0 + i < n_bytes) w |= (uint32_t)p.allow[byte0 + i] << (8 * i);
                        valid &= w;
                    }
                    m &= valid;
                    if (__reduce_or_sync(0xffffffffu, m) != 0u) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = (m >> j) & 1u ? __uint_as_float(v[j]) : -INFINITY;
#pragma unroll
                            for (int i = 0; i < kSampleM; ++i) {     // sorted insert: bubble x down the list
                                const float hi = fmaxf(best[i], x);
                                x = fminf(best[i], x);
                                best[i] = hi;
                            }
                        }
                    }
                    continue;
                }
                uint32_t um = __reduce_or_sync(0xffffffffu, m);
                while (um) {
                    const int j = __ffs(um) - 1;
                    um &= um - 1;
                    const float s = __uint_as_float(tc_ld1(tbase + c * 32 + j));
                    tc_wait_ld();
                    if ((m >> j) & 1u) {
                        const uint32_t row = row0 + c * 32 + j;
                        if (row < (uint32_t)p.n_rows && bitmap_test(p.allow, row)) {
                            if (MODE == 0) {
                                float x = s;         // sorted insert: bubble x down the list
#pragma unroll
                                for (int i = 0; i < kSampleM; ++i) {
                                    const float hi = fmaxf(best[i], x);
                                    x = fminf(best[i], x);
                                    best[i] = hi;
                                }
                            } else {
                                s_pend[npend++] = make_key(s, row);
                                if (npend == kPend) {                       // rare: many survivors in one tile
                                    const int slot = atomicAdd(my_cnt, kPend);
                                    for (int i = 0; i < kPend; ++i)
                                        if (slot + i < p.list_cap) my_list[slot + i] = s_pend[i];
                                    npend = 0;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NCTA == 2) mbar_arrive_remote(buf ? tempty_leader1 : tempty_leader0);
                else mbar_arrive(&tempty_bar[buf]);
            }
            if (MODE == 1 && npend > 0) {              // counts past the capacity flag an overflow
                const int slot = atomicAdd(my_cnt, npend);
                for (int i = 0; i < npend; ++i)
                    if (slot + i < p.list_cap) my_list[slot + i] = s_pend[i];
            }
        }
        if (MODE == 0 && cur_qb >= 0) flush_sample(cur_qb);
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();     // the peer is done with this CTA's barriers / operands / TMEM
    if (warp == 1) {
        tc_fence_after();
        if (NCTA == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- query preparation: fp32 -> bf16 (RNE) + exact norm of the rounding residual ---------------------------
template <typename T16>
__global__ void query_prep_kernel(const float* __restrict__ q, int n_queries, int n_padded, int dim,
                                  T16* __restrict__ q16, float* __restrict__ resid_norm) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_padded) return;
    double r2 = 0.0;
    for (int c = lane; c < dim; c += 32) {
        float v = w < n_queries ? q[(size_t)w * dim + c] : 0.f;
        T16 h;
        float back;
        if constexpr (sizeof(T16) == sizeof(__nv_bfloat16) && std::is_same<T16, __nv_bfloat16>::value) {
            h = __float2bfloat16_rn(v);
            back = __bfloat162float(h);
        } else {
            h = __float2half_rn(v);
            back = __half2float(h);
        }
        q16[(size_t)w * dim + c] = h;
        double d = (double)v - (double)back;
        r2 += d * d;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    if (lane == 0 && w < n_queries) resid_norm[w] = (float)(sqrt(r2) * 1.0000001 + 1e-30);
}

cudaError_t query_prep_launch(const float* q, int n_queries, int n_padded, int dim, void* q16, float* resid_norm,
                              int fp16, cudaStream_t st) {
    const int warps_per_block = 8;
    int grid = (n_padded + warps_per_block - 1) / warps_per_block;
    if (fp16)
        query_prep_kernel<__half><<<grid, warps_per_block * 32, 0, st>>>(q, n_queries, n_padded, dim,
                                                                        reinterpret_cast<__half*>(q16), resid_norm);
    else
        query_prep_kernel<__nv_bfloat16><<<grid, warps_per_block * 32, 0, st>>>(
            q, n_queries, n_padded, dim, reinterpret_cast<__nv_bfloat16*>(q16), resid_norm);
    return cudaGetLastError();
}

// bf16 shadow of an fp32 / fp16 corpus + max over rows of ||x - bf16(x)||_2
template <int DT>
__global__ void shadow_kernel(const void* __restrict__ rows, int64_t n_rows, int dim, __nv_bfloat16* __restrict__ out,
                              float* __restrict__ max_resid) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        float r2 = 0.f;
        for (int c = lane; c < dim; c += 32) {
            float v;
            if constexpr (DT == RAG_F32) v = reinterpret_cast<const float*>(rows)[(size_t)r * dim + c];
            else v = __half2float(reinterpret_cast<const __half*>(rows)[(size_t)r * dim + c]);
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            out[(size_t)r * dim + c] = h;
            float d = v - __bfloat162float(h);          // exact in fp32
            r2 = fmaf(d, d, r2);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        best = fmaxf(best, r2);
    }
    if (lane == 0 && best > 0.f)
        atomicMax(reinterpret_cast<int*>(max_resid), __float_as_int(sqrtf(best) * 1.00001f));
}

cudaError_t shadow_launch(const void* rows, int dtype, int64_t n_rows, int dim, void* out, float* max_resid,
                          cudaStream_t st) {
    int64_t g = (n_rows * 32 + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    if (dtype == RAG_F32)
        shadow_kernel<RAG_F32><<<(int)g, 256, 0, st>>>(rows, n_rows, dim, reinterpret_cast<__nv_bfloat16*>(out), max_resid);
    else
        shadow_kernel<RAG_F16><<<(int)g, 256, 0, st>>>(rows, n_rows, dim, reinterpret_cast<__nv_bfloat16*>(out), max_resid);
    return cudaGetLastError();
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major [rows][dim] bf16 matrix, box = box_rows x 64 elements, 128B swizzle, OOB rows read as zero
static bool make_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int box_rows, int fp16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    cuuint32_t box[2] = {(cuuint32_t)GT_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int gemm_padded_queries(int n_queries) { return (n_queries + GT_M - 1) / GT_M * GT_M; }

int g_balance_tail = 1;   // split the leftover tiles of the main pass by items (option "balance_tail")
void gemm_set_balance_tail(int v) { g_balance_tail = v != 0; }
int g_pair_mode = 1;      // main pass as CTA pairs (tcgen05 cta_group::2) when the query blocks pair up (option "pair_mode")
void gemm_set_pair_mode(int v) { g_pair_mode = v != 0; }
int g_sample_resident = 1;   // small samples keep their rows resident in shared memory (option "sample_resident")
void gemm_set_sample_resident(int v) { g_sample_resident = v != 0; }
int g_sample_div = 1;     // multiplies the survivor target of the sample pass (option "sample_div", experiments)
void gemm_set_sample_div(int v) { g_sample_div = v < 1 ? 1 : (v > 8 ? 8 : v); }
int gemm_sample_m() { return kSampleM; }
int gemm_max_batch() { return kMaxQBlocks * GT_M; }

size_t gemm_plan(GemmParams& p, int sm_count, int smem_limit, int* grid_out) {
    p.n_qblocks = gemm_padded_queries(p.n_queries) / GT_M;
    if (p.n_qblocks > kMaxQBlocks) return 0;
    // barriers, tau (query blocks rounded up to a pair), parked keys
    const size_t tail = 256 + (size_t)((p.n_qblocks + 1) / 2 * 2) * GT_M * 4 + (size_t)GT_M * kPend * 8;
    int stages = (int)((smem_limit - 1024 - (long)tail) / GT_STAGE_BYTES);
    if (stages > 4) stages = 4;
    if (stages < 2) return 0;
    p.n_stages = stages;
    p.ring_bytes = stages * GT_STAGE_BYTES;
    const int64_t n_tiles = (p.n_rows + GT_N - 1) / GT_N;
    int grid = (int)(n_tiles < sm_count ? (n_tiles > 0 ? n_tiles : 1) : sm_count);
    *grid_out = grid;
    p.n_lists = grid;
    p.balance_tail = g_balance_tail;
    // pair mode (main pass): 32 KB stages, one CTA pair per row tile; only when the query blocks pair up
    // (an odd block count would leave one CTA of the last pair multiplying padding) and every pair has work
    {
        const int pair_stage = GT_A_BYTES + GT_B_BYTES / 2;
        int ps = (int)((smem_limit - 1024 - (long)tail) / pair_stage);
        if (ps > 6) ps = 6;
        p.n_stages_pair = ps;
        p.pair = g_pair_mode && p.n_qblocks % 2 == 0 && ps >= 2 && n_tiles >= sm_count && sm_count >= 2;
        p.smem_pair = (size_t)ps * pair_stage + tail + 1024;
        p.ring_bytes_pair = ps * pair_stage;
    }
    // Sample pass: the m-th best (m = kSampleM = 8) of a sample that holds a fraction f of the rows lets
    // ~(1/f) * Gamma(m) rows per query through the main pass.  Every survivor costs the main pass a slow-path
    // visit, a larger sample costs the sample pass: aim for E = max(1024, 4*kp) survivors per query
    // (measured sweet spot for k = 10 and k = 100), i.e. f = 8 / E.  Fewer than ~1.5*k survivors (which
    // would flag the query for the fallback pass) then has probability P(Gamma(8) < 12*k/E) < 1e-4.
    // Column-granular.
    const int64_t tiles_per_cta = (n_tiles + grid - 1) / grid;
    const int64_t rows_per_cta = tiles_per_cta * GT_N;
    const int64_t e_target = (int64_t)(4 * p.kp > 1024 ? 4 * p.kp : 1024) * (g_sample_div > 0 ? g_sample_div : 1);
    {
        p.use_sample = 1;
        int64_t want = (rows_per_cta * kSampleM + e_target - 1) / e_target;   // sample rows per CTA
        if (want < 2) want = 2;
        if (want <= GT_N) {
            p.sample_tiles = 1;
            p.sample_chunks = (int)((want + 31) / 32);
            const int last = (int)(want - (int64_t)(p.sample_chunks - 1) * 32);
            p.sample_last_mask = last >= 32 ? 0xffffffffu : ((1u << last) - 1u);
            p.sample_step = 1;
        } else {
            p.sample_tiles = (int)((want + GT_N - 1) / GT_N);
            const int per_tile = (int)((want + p.sample_tiles - 1) / p.sample_tiles);   // columns of every sample tile
            p.sample_chunks = (per_tile + 31) / 32;
            const int last = per_tile - (p.sample_chunks - 1) * 32;
            p.sample_last_mask = last >= 32 ? 0xffffffffu : ((1u << last) - 1u);
            p.sample_step = (int)(tiles_per_cta / p.sample_tiles > 0 ? tiles_per_cta / p.sample_tiles : 1);
        }
    }
    // small sample (one tile, <= 64 columns): the sample rows stay resident in shared memory (dim/64 x 8 KB) and
    // the MMA is 64 columns wide; the queries stream through a ring of 16 KB stages
    p.sample_resident = 0;
    if (g_sample_resident && p.sample_tiles == 1 && p.sample_chunks <= 2) {
        const long res = (long)(p.dim / GT_K) * 64 * GT_K * 2;
        int sa = (int)((smem_limit - 1024 - (long)tail - res) / GT_A_BYTES);
        if (sa > 6) sa = 6;
        if (sa >= 3) {
            p.sample_resident = 1;
            p.n_stages_sample = sa;
            p.ring_bytes_sample = (int)(res + (long)sa * GT_A_BYTES);
            p.smem_sample = (size_t)p.ring_bytes_sample + tail + 1024;
        }
    }
    // one candidate list per query, filled by all CTAs: E survivors expected, 4x head-room
    p.list_cap = (int)(4 * e_target);
    return (size_t)stages * GT_STAGE_BYTES + tail + 1024;   // + slack for the 1024-byte alignment of the ring
}

// main pass as CTA pairs: cluster launch, one pair per SM pair that can be co-resident
static cudaError_t gemm_launch_pair(const GemmParams& p, const void* q16, const void* x16, int sm_count, cudaStream_t st) {
    CUtensorMap map_q, map_x;
    if (!make_map(&map_q, q16, gemm_padded_queries(p.n_queries), p.dim, GT_M, p.fp16_operands))
        return cudaErrorNotSupported;
    if (!make_map(&map_x, x16, p.n_rows, p.dim, GT_N / 2, p.fp16_operands)) return cudaErrorNotSupported;
    auto kern = dense_gemm_topk_kernel<1, 2>;
    cudaError_t e = ensure_dynamic_smem_of(kern, (size_t)(p.smem_pair));
    if (e != cudaSuccess) return e;
    GemmParams pp = p;
    pp.n_stages = p.n_stages_pair;
    pp.ring_bytes = p.ring_bytes_pair;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(GT_THREADS);
    cfg.dynamicSmemBytes = p.smem_pair;
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static int max_pairs = -1;            // CTA pairs that are resident together (persistent kernel: one wave)
    if (max_pairs < 0) {
        cfg.gridDim = dim3(sm_count / 2 * 2);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) return e;
        max_pairs = n;
    }
    const int64_t n_tiles = (p.n_rows + GT_N - 1) / GT_N;
    int pairs = max_pairs < sm_count / 2 ? max_pairs : sm_count / 2;
    if (n_tiles < pairs) pairs = (int)n_tiles;
    if (pairs < 1) return cudaErrorLaunchOutOfResources;
    cfg.gridDim = dim3(2 * pairs);
    return cudaLaunchKernelEx(&cfg, kern, map_q, map_x, pp);
}

cudaError_t gemm_launch(const GemmParams& p, int mode, const void* q16, const void* x16, int grid, size_t smem,
                        cudaStream_t st) {
    if (mode == 1 && p.pair) return gemm_launch_pair(p, q16, x16, grid, st);
    const bool resident = mode == 0 && p.sample_resident;
    GemmParams pp = p;
    if (resident) {
        pp.n_stages = p.n_stages_sample;
        pp.ring_bytes = p.ring_bytes_sample;
        smem = p.smem_sample;
    } else {
        pp.sample_resident = 0;
    }
    CUtensorMap map_q, map_x;
    if (!make_map(&map_q, q16, gemm_padded_queries(p.n_queries), p.dim, GT_M, p.fp16_operands))
        return cudaErrorNotSupported;
    if (!make_map(&map_x, x16, p.n_rows, p.dim, resident ? 64 : GT_N, p.fp16_operands)) return cudaErrorNotSupported;
    auto kern = mode == 0 ? dense_gemm_topk_kernel<0, 1> : dense_gemm_topk_kernel<1, 1>;
    cudaError_t e = ensure_dynamic_smem_of(kern, (size_t)(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, GT_THREADS, smem, st>>>(map_q, map_x, pp);
    return cudaGetLastError();
}

}  // namespace b200rag
