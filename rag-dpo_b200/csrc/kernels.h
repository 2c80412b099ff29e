// kernels.h — host-visible launchers of the b200rag kernels (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200rag.h"

namespace b200rag {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size) instead of before every launch: the
// attribute belongs to the current device's copy of the function, and the call is not free on a latency path
// (a single-query search is three or four launches).  Thread-safe; defined in corpus.cu.
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);
template <typename K>
inline cudaError_t ensure_dynamic_smem_of(K kernel, size_t bytes) {
    return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}

// ---- dense_scan.cu ---------------------------------------------------------
struct ScanParams {
    const void* rows;        // n_rows x dim, storage dtype, row-major
    const float* q;          // n_queries x dim fp32 (device)
    const uint8_t* allow;    // nullable row bitmap (device)
    int64_t n_rows;
    int dim;
    int n_queries;           // 1..4 in this launch
    int nq_t;                // kernel template width (1, 2 or 4) >= n_queries; fixes the warp count
    int kp;                  // per-list capacity (power of two >= 16)
    int mode;                // 0 = running top-k, 1 = collect >= tau
    // derived by scan_plan
    int row_bytes, tile_rows, tile_bytes, n_stages, n_lists;
    // mode 0 output: n_queries x n_lists x kp keys
    uint64_t* cand;
    // mode 1
    const float* tau;        // per-query filter-score threshold
    unsigned* collect_count; // per query
    uint32_t* collect_rows;  // n_queries x collect_cap
    int collect_cap;
    // device-driven fallback (stream-ordered calls): the launch serves queries [active_first, *n_active_dev) of
    // the compacted fallback list (at most the template width) and returns at once when there are none
    const int* n_active_dev;
    int active_first;
};
int scan_nq_template(int n_queries);
size_t scan_plan(ScanParams& p, int dtype, int sm_count, int smem_limit, int* grid_out, int* nch_out);
cudaError_t scan_launch(const ScanParams& p, int dtype, int nch, int grid, size_t smem, cudaStream_t st);

// ---- dense_gemm.cu ---------------------------------------------------------
struct GemmParams {
    int64_t n_rows;
    int dim;
    int n_queries;           // B (the bf16 query buffer is padded to a multiple of 128 rows)
    int kp;                  // candidates kept per query after the merge (power of two >= 64)
    const uint8_t* allow;    // nullable row bitmap (device)
    // sample pass (mode 0): padded_B x n_lists x m keys out (m = gemm_sample_m())
    uint64_t* sample_keys;
    // main pass (mode 1): tau_keys = merged sample (padded_B x m, sorted desc; [m-1] is the threshold)
    const uint64_t* tau_keys;
    uint64_t* cand;          // padded_B x list_cap keys: one list per query, appended to by every CTA
    int32_t* cand_cnt;       // padded_B survivors per query (may exceed list_cap = overflow), zeroed by the caller
    // derived by gemm_plan
    int list_cap;            // entries per query list
    int use_sample;          // always 1 (kept for the launcher)
    int n_lists, n_stages, n_qblocks, sample_tiles, sample_step, sample_chunks;
    uint32_t sample_last_mask;   // columns of the last sample chunk that count
    int balance_tail;            // main pass: split the leftover tiles by (tile, query block) items
    int fp16_operands;           // 1: both operands are fp16 (fp16 corpus used as stored), 0: bf16
    // pair mode of the main pass (tcgen05 cta_group::2): decided by gemm_plan
    int pair, n_stages_pair, ring_bytes_pair;
    size_t smem_pair;
    int ring_bytes;              // bytes of the operand ring in front of the barriers (set per launch)
    // resident sample (small samples): decided by gemm_plan
    int sample_resident, n_stages_sample, ring_bytes_sample;
    size_t smem_sample;
};
int gemm_sample_m();
void gemm_set_sample_div(int v);
void gemm_set_balance_tail(int v);
void gemm_set_pair_mode(int v);
void gemm_set_sample_resident(int v);
int gemm_max_batch();
int gemm_padded_queries(int n_queries);
size_t gemm_plan(GemmParams& p, int sm_count, int smem_limit, int* grid_out);
cudaError_t gemm_launch(const GemmParams& p, int mode, const void* q16, const void* x16, int grid, size_t smem,
                        cudaStream_t st);
// fp32 queries -> bf16 (padded rows zeroed) + ||q - bf16(q)||_2 per query
cudaError_t query_prep_launch(const float* q, int n_queries, int n_padded, int dim, void* q16, float* resid_norm,
                              int fp16, cudaStream_t st);
// bf16 shadow of an fp32/fp16 corpus + max_r ||x_r - bf16(x_r)||_2 (atomicMax into *max_resid)
cudaError_t shadow_launch(const void* rows, int dtype, int64_t n_rows, int dim, void* out, float* max_resid,
                          cudaStream_t st);

// ---- dense_select.cu -------------------------------------------------------
// sample threshold: the m (= 8) best of each query's n_keys sample keys -> top (B x m, descending)
cudaError_t sample_tau_launch(const uint64_t* keys, int n_queries, int n_keys, int m, uint64_t* top, cudaStream_t st);
// refine: canonical fp64 score of every candidate in top, sort by (score desc,
// row asc), write the first k, and raise flag[b] when the margin check fails.
struct RefineParams {
    const uint64_t* top;     // B x kp
    const void* rows;
    const float* q;          // B x dim
    int dtype, dim, kp, k, B;
    // filter error bound: eps = (eps_rel*|q| + q_resid[b]) * max_row_norm + 1.004*|q| * x_resid
    double eps_rel;          // accumulation error / (|q| * max|x|)
    const float* q_resid;    // nullable, per query: ||q - bf16(q)||_2 (tensor-core path)
    const float* x_resid;    // nullable device scalar: max_r ||x_r - bf16(x_r)||_2 (bf16 shadow)
    const float* max_row_norm;  // device scalar
    // threshold-capture mode (tensor-core path): rows outside the candidate lists have filter score
    // < tau_q (tau_keys[b*tau_stride + tau_stride-1], 0 = none); an overflowed list (merge_refine_launch's
    // `overflow`) voids that guarantee
    const uint64_t* tau_keys;
    int tau_stride;
    int32_t* out_rows;       // B x k
    double* out_scores;      // B x k
    int32_t* out_counts;     // B
    int32_t* flags;          // B: 1 = margin check failed -> fallback pass needed
    float* tau;              // B: threshold for the fallback collect pass
    int32_t* n_flagged;      // scalar counter
    // sharded corpora (one process, several GPUs): the results go straight into the primary device's gather
    // buffer as GLOBAL ids of the block-cyclic layout ((l / block * n_shards + shard) * block + l % block)
    int64_t* out_gids;       // nullable: B x k, written instead of out_rows
    int shard, n_shards, shard_block;
    // device-driven fallback (stream-ordered calls, nullable): the last CTA to finish compacts the flagged queries
    // into fb_index / fb_tau / fb_q (at most fb_max; the surplus gets out_counts = -1) and zeroes the collect counters
    unsigned* fb_done;       // CTA completion counter (zero between launches)
    int32_t* fb_n;           // number of compacted queries
    int32_t* fb_index;       // their query numbers
    float* fb_tau;           // their collect thresholds
    float* fb_q;             // their query vectors (fb_max x dim)
    unsigned* fb_counts;     // collect counters (zeroed)
    int fb_max;
};
__host__ __device__ inline int64_t shard_global_row(int64_t local, int shard, int n_shards, int block) {
    return ((local / block) * n_shards + shard) * block + local % block;
}
cudaError_t refine_launch(const RefineParams& p, cudaStream_t st);
// merge + refine fused: cand = B x n_lists x list_len keys (any order, 0 = empty); counts: NULL (all entries
// valid), or per (query, list) valid entries, or — flat_counts != 0 — ONE count per query for the whole
// contiguous n_lists*list_len block; overflow (nullable): set to 1 for queries whose count exceeds the
// capacity.  The merged top-kp never leaves shared memory (p.top is not used).
// sorted_lists != 0: every list is sorted descending with its empties last (the scan kernel's lists)
cudaError_t merge_refine_launch(const uint64_t* cand, const int32_t* counts, int flat_counts, int n_lists, int list_len,
                                int sorted_lists, int32_t* overflow, const RefineParams& p, cudaStream_t st);
// fallback tail: exact scores of the collected rows + top-k select, one CTA per query
struct CollectSelectParams {
    const uint32_t* rows_list;   // nq x cap
    const unsigned* counts;      // nq
    int cap;
    const int32_t* query_index;  // nq: which output slot each collected query maps to
    const void* rows;
    const float* q;              // nq x dim (the flagged queries, compacted)
    int dtype, dim, k, nq;
    double* scratch_scores;      // nq x cap
    int32_t* out_rows;           // B x k (indexed through query_index)
    double* out_scores;
    int32_t* out_counts;
    int64_t* out_gids;           // nullable: global ids instead of out_rows (sharded corpora)
    int shard, n_shards, shard_block;
    const int* n_active_dev;     // nullable: device-side number of queries (CTAs beyond it return at once)
};
cudaError_t collect_select_launch(const CollectSelectParams& p, cudaStream_t st);
// G sorted (score,id) lists per query -> global top-k
cudaError_t merge_exact_launch(const double* scores, const int64_t* ids, int G, int B, int k, int64_t rank_stride,
                               double* out_scores, int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

// ---- exchange.cu -----------------------------------------------------------
constexpr int kMaxExchangeRanks = 16;
struct ExchangeDev {
    uint8_t* peer_base[kMaxExchangeRanks];     // every rank's payload area (mine included), as mapped in this process
    uint64_t* peer_flags[kMaxExchangeRanks];   // every rank's flag area
    uint8_t* my_base;
    uint64_t* my_flags;
    unsigned* done_counter;                     // local: blocks of the push kernel that have finished; [1] = error word
    long long timeout_cycles;                   // bound of the flag wait (option "exchange_timeout_ms")
    int world, rank;
    size_t slot_bytes;                          // payload capacity per (parity, source rank)
};
// push my (scores | ids) block into every rank's buffer and flag the epoch (ids: int64 global ids, or — my_rows
// != NULL — local int32 rows turned into global ids row_lo + row on the way); then wait for all ranks, merge -> top-k
cudaError_t exchange_push_launch(const ExchangeDev& ex, const double* my_scores, const int64_t* my_ids,
                                 const int32_t* my_rows, int64_t row_lo, const int32_t* my_counts, int B, int k,
                                 uint64_t epoch, cudaStream_t st);
cudaError_t exchange_merge_launch(const ExchangeDev& ex, int B, int k, uint64_t epoch, double* out_scores,
                                  int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

// ---- corpus.cu -------------------------------------------------------------
cudaError_t convert_rows_launch(const float* src, void* dst, int dtype, int64_t n_elems, cudaStream_t st);
cudaError_t widen_rows_launch(const void* src, int dtype, float* dst, int64_t n_elems, cudaStream_t st);
cudaError_t row_norm_max_launch(const void* rows, int dtype, int64_t n_rows, int dim, float* max_norm,
                                cudaStream_t st);
cudaError_t fill_synthetic_launch(void* rows, int dtype, int64_t row0, int64_t n_rows, int dim, uint64_t seed,
                                  int64_t gen_row0, cudaStream_t st);
cudaError_t gather_rows_launch(const void* src, void* dst, const int64_t* keep, int64_t nkeep, int row_bytes,
                               cudaStream_t st);

// ---- bm25.cu ---------------------------------------------------------------
constexpr int kBm25Range = 4096;          // rows per CTA of the exact range kernel
constexpr int kBm25Block = 16384;         // rows per block of the filter index (14-bit local row, range tables, columns)
struct Bm25Device {
    int64_t n_docs, n_terms, nnz;
    int64_t* term_ptr;      // n_terms + 1
    int32_t* post_row;      // nnz, ascending per term
    double* post_impact;    // nnz: tf*(k1+1) / (tf + k1*(1-b+b*dl/avgdl))
    double* idf;            // n_terms
    double* score;          // n_docs accumulator, all-zero between queries (rag_bm25_scores)
    // ---- filter index of the fast path (built by bm25_index_build) ----
    uint32_t* post_pack;    // nnz + 4: (row & 16383) << 18 | q, q = idf*impact in units of `unit`, rounded up (+1)
    int2* term_info;        // n_terms: x = class << 30 | slot (rng_off row), y = column (dense_col) of a DENSE term
    int32_t* rng_off;       // n_tabled x (n_blocks + 1): first posting with row >= b * 16384, relative to term_ptr[t]
    uint16_t* dense_col;    // n_dense x (n_blocks * 16384): ceil(q / 4) of the term's posting on that row, 0 = no posting
    int n_blocks;
    int n_dense;            // dense columns
    int fast_ok;            // 1: all idf >= 0 and the packed stream exists -> the integer filter bound is valid
};
// term classes (term_info.x >> 30)
constexpr int kBmLow = 0;    // short list, no table: every CTA scans the whole list and keeps the rows of its range
constexpr int kBmMid = 1;    // rng_off row gives the run of the CTA's range inside the packed stream
constexpr int kBmDense = 2;  // rng_off row (exact recompute) + a dense 16-bit column: coalesced, accumulated in registers
constexpr int kBmSkip = 3;   // idf == 0 or empty list: contributes nothing
constexpr int kBmDenseShift = 2;     // a column entry is ceil(q / 2^2): 16 bits, at most 3 + 2 units above the product
cudaError_t bm25_impact_launch(const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, int64_t nnz,
                               double avgdl, double k1, double b, double* impact, cudaStream_t st);
// max over the postings of idf[t] * impact[p] (atomicMax on the bits of a positive double; *cmax zeroed by the caller)
cudaError_t bm25_cmax_launch(const Bm25Device& ix, unsigned long long* cmax, cudaStream_t st);
// packed stream + range tables + dense columns (tabled_terms: MID and DENSE terms in slot order, dense_terms: in
// column order; the columns must be zeroed by the caller)
cudaError_t bm25_index_build_launch(const Bm25Device& ix, double unit, const int32_t* tabled_terms, int n_tabled,
                                    const int32_t* dense_terms, int n_dense, cudaStream_t st);
// one token of one query: score[row] += w * impact[p] over postings [lo, hi)
cudaError_t bm25_accumulate_launch(const Bm25Device& ix, int64_t lo, int64_t hi, double w, cudaStream_t st);
int bm25_harvest_grid(int64_t total_postings, int sm_count);     // grid for the reset pass
cudaError_t bm25_reset_launch(const Bm25Device& ix, const int64_t* d_ranges, int n_ranges, int grid, cudaStream_t st);
size_t bm25_key_bytes();
// robust path: grid (row ranges of 4096, queries); fp64 accumulators in shared memory, tokens in order
int bm25_range_lists(int64_t n_docs);
cudaError_t bm25_range_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr,
                              const int32_t* d_q_index, int Q, const uint8_t* allow, int kp, int k, void* cand,
                              int32_t* out_rows, double* out_scores, int32_t* out_counts, cudaStream_t st);
// fast path: integer filter pass over the packed postings (range heads) -> threshold, survivors, exact fp64
// recompute, final order (counts = -1: redo on the robust path)
bool bm25_fast_supported(const Bm25Device& ix, int k, int max_query_tokens);
size_t bm25_fast_scratch_bytes(const Bm25Device& ix, int k, int Q, int max_query_tokens);
int bm25_fast_chunk(const Bm25Device& ix, int max_query_tokens);     // queries per bm25_fast_launch
cudaError_t bm25_fast_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr, int q0, int Q,
                             int max_query_tokens, const uint8_t* allow, int k, void* scratch, int32_t* out_rows,
                             double* out_scores, int32_t* out_counts, cudaStream_t st);
// selective row filter: exact scores of the listed rows only (no posting stream at all), top-k
constexpr int kBm25MaxListedRows = 4096;
cudaError_t bm25_rows_launch(const Bm25Device& ix, const int32_t* d_q_terms, const int32_t* d_q_ptr, int Q,
                             const int32_t* d_rows, int n_rows, int k, int32_t* out_rows, double* out_scores,
                             int32_t* out_counts, cudaStream_t st);

// ---- rrf.cu ----------------------------------------------------------------
int rrf_max_entries();
cudaError_t rrf_launch(const int32_t* ids, const double* weights, int Q, int R, int L, int rrf_k, int top,
                       int32_t* out_ids, double* out_scores, int32_t* out_counts, cudaStream_t st);

int rerank_max_candidates();
cudaError_t rerank_select_launch(const float* scores, const double* boosts, const int32_t* lens, int Q, int L, int top_k,
                                 double min_score, int32_t* out_idx, double* out_scores, int32_t* out_counts,
                                 cudaStream_t st);

}  // namespace b200rag
