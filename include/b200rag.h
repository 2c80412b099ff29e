/* b200rag.h — C ABI of libb200rag.so: the B200-native (sm_100a) retrieval hot
 * path that stands behind RAG-DPO's injected `collection` / `chunk_bm25_index`
 * objects.  The reference has NO FFI of its own (pure Python, duck typing);
 * each entry point below names the reference call it replaces (paths relative
 * to /root/reference).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative RAG_E* code; the text
 *     of the last failure on the calling thread is rag_last_error().
 *   - "host" pointers are plain host memory owned by the caller; "_dev"
 *     entry points take device pointers (HBM-resident inputs / torch interop).
 *   - the library owns all device memory behind its handles.  Host-pointer
 *     calls are synchronous (they return after their stream work finished);
 *     "_dev" calls are STREAM-ORDERED (they only queue work, see each one).
 *     Every handle has its own stream, scratch and lock: calls on different
 *     handles (dense, BM25, RRF) do not serialise on each other.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with RAG_ENODEV.
 *   - dense scores are the CANONICAL fp64 scores of DESIGN.md §3 (fixed-order
 *     fp64 sum of exact products), so ids and scores are bit-identical to the
 *     oracle.  Order: score descending, ties -> lowest row.
 *   - one process can drive several GPUs (rag_init_devices): a SHARDED corpus /
 *     BM25 index spreads its rows over the shard slots in blocks of
 *     RAG_SHARD_BLOCK rows (block-cyclic), every call addresses GLOBAL rows, and
 *     the per-shard top-k lists are merged on the first slot's device over
 *     NVLink peer memory.  The Python boundary stays one synchronous call.
 */
#ifndef B200RAG_H
#define B200RAG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAG_OK 0
#define RAG_EINVAL (-1)   /* bad argument */
#define RAG_ENODEV (-2)   /* no usable sm_100 device / CUDA error */
#define RAG_ENOMEM (-3)
#define RAG_ERANGE (-4)   /* k, dim or batch outside the supported range */
#define RAG_ECUDA (-5)

/* storage dtype of corpus rows */
#define RAG_F32 0
#define RAG_BF16 1
#define RAG_F16 2

#define RAG_MAX_K 224         /* largest n_results / top_k served by one fused select */
#define RAG_SHARD_BLOCK 1024  /* rows per block of the block-cyclic shard layout */
#define RAG_MAX_COLUMNS 32    /* coded metadata columns per corpus / index */

typedef struct rag_corpus rag_corpus_t;
typedef struct rag_bm25 rag_bm25_t;

/* ---- runtime ------------------------------------------------------------ */
int rag_init(int device);                 /* bind the process to one GPU (one process per GPU, e.g. under torchrun) */
/* one process, several GPUs: shard slot i of a sharded handle runs on devices[i] (a device may repeat: several
 * shards per GPU); peer access is enabled between all of them.  devices[0] is the primary device (merges). */
int rag_init_devices(int n_slots, const int* devices);
int rag_slot_count(int* n_slots);
int rag_set_stream(void* cuda_stream);    /* run the primary device's work on the caller's stream (e.g. torch's); NULL = own */
const char* rag_last_error(void);
int rag_abi_version(void);
int rag_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes);
/* runtime knobs: "tc_min_batch" (smallest batch served by the tcgen05 contraction path, default 2),
 * "tc_b1_shadow" (1 = a single query on an fp32 corpus is filtered through the bf16 shadow, default 1),
 * "pair_mode" (1 = main pass as CTA pairs with tcgen05 cta_group::2 when the query blocks pair up, default 1),
 * "sample_resident", "balance_tail", "sample_div" (experiments; defaults 1),
 * "exchange_timeout_ms" (bound of the exchange's flag wait, default 10000),
 * "bm25_fast" (1 = packed-postings filter path, default 1), "bm25_rows_max" (largest row filter served by the
 * listed-rows kernel, default 4096), "bm25_dense_div" (terms with df >= n_docs / div get a dense 16-bit column when
 * an index is built, default 8; 0 = none), "bm25_tile" (force 1 / 2 / 4 chunks of 4096 rows per filter CTA,
 * default 0 = by launch size), "bm25_tma" (1 = block tiles through the kernel that stages the posting runs through
 * shared memory with TMA bulk copies, default 1), "bm25_acc16" (1 = that kernel keeps 16-bit accumulators in a
 * coarser unit: 5 CTAs per SM, default 1), "bm25_by_block" (1 = its CTAs in (block, query) order, default 1), "bm25_inline_resolve" (1 = launches of <= 4
 * queries resolve their token records inside the filter kernel instead of a separate launch, default 1) */
int rag_set_option(const char* key, int64_t value);
/* page-locked host memory: buffers allocated here are DMA'd directly by the host-pointer entry points
 * (no staging copy); any other host pointer is staged through an internal pinned block. */
int rag_host_alloc(void** out, size_t bytes);
int rag_host_free(void* p);
/* per-stage time (ms) of the last dense / bm25 call.  CUDA events on the launching stream: [0] filter stage
 * (scan kernel, or query prep + sample + threshold + contraction; BM25: the whole search), [1] unused gap,
 * [2] select + fp64 refine, [3] fallback pass (0 if not taken), [6] the contraction's main pass alone.
 * Host clock, host-buffer dense call only: [4] time spent queueing work before the one synchronisation,
 * [5] whole call. */
int rag_last_timings(float* ms, int n);
/* counters since rag_init: [0] kernels launched, [1] fallback passes taken (host-checked calls), [2] queries
 * those passes re-did */
int rag_counters(int64_t* out, int n);
/* Test hook (no reference counterpart): the candidate lists the tensor-core filter of the last dense call on a
 * single-shard corpus left behind. Keys: (order-preserving image of the fp32 filter score) << 32 | ~row; out_counts[b]
 * = entries written for query b (-1: overflowed list); *eps_rel = the accumulation error bound of the margin check
 * relative to |q| * max|x|. tests/test_gpu.py measures |filter - exact| against that bound with it. */
int rag_debug_last_candidates(const rag_corpus_t* c, int n_queries, uint64_t* out_keys, int64_t cap_per_query,
                              int32_t* out_counts, double* eps_rel);

/* ---- corpus: the chunk-embedding matrix -----------------------------------
 * replaces the vector segment behind chromadb's Collection: written through
 * collection.add (src/processing/create_chromadb_index.py:374-379,
 * src/processing/ingest_enterprise.py:241-246). Rows are appended in order;
 * row index == insertion order. dim % 64 == 0, dim <= 2048 (fp32: <= 1024). */
int rag_corpus_create(rag_corpus_t** out, int64_t capacity_rows, int dim, int dtype);
/* rows spread block-cyclically over the first n_shards shard slots (rag_init_devices) */
int rag_corpus_create_sharded(rag_corpus_t** out, int64_t capacity_rows, int dim, int dtype, int n_shards);
int rag_corpus_destroy(rag_corpus_t* c);
int rag_corpus_reserve(rag_corpus_t* c, int64_t capacity_rows);
/* append nrows fp32 host rows at row0 (== current count, or overwrite below it);
 * rows are converted to the storage dtype on the device (round-to-nearest-even). */
int rag_corpus_upload(rag_corpus_t* c, int64_t row0, int64_t nrows, const float* host_rows);
/* stored values widened to fp32 (collection.get(include=["embeddings"])) */
int rag_corpus_download(const rag_corpus_t* c, int64_t row0, int64_t nrows, float* host_rows);
/* collection.delete (src/processing/ingest_enterprise.py:272,304): the rows become tombstones — a device bitmap
 * every search tests before its top-k; O(n) in the rows deleted, row numbers do not move */
int rag_corpus_delete_rows(rag_corpus_t* c, const int64_t* rows, int64_t n);
/* physically keep only the listed rows, in the listed (ascending) order, and drop all tombstones (single-shard
 * corpora; the caller renumbers its row-indexed lists the same way) */
int rag_corpus_compact(rag_corpus_t* c, const int64_t* keep_rows, int64_t nkeep);
int rag_corpus_count(const rag_corpus_t* c, int64_t* n);            /* rows incl. tombstones */
int rag_corpus_live_count(const rag_corpus_t* c, int64_t* n_live);
/* coded metadata column `column` (< RAG_MAX_COLUMNS) of rows [row0, row0+nrows): one int32 code per row, -1 = the
 * row's metadata has no such key (collection.add / collection.update, tag_all_chunks.py:215).  The value ->
 * code dictionaries stay with the caller. */
int rag_corpus_set_codes(rag_corpus_t* c, int column, int64_t row0, int64_t nrows, const int32_t* codes);
/* deterministic synthetic unit rows generated on the device (bench/test input:
 * counter-based hash of (seed, gen_row0 + i, col) -> uniform(-1,1) -> L2-normalised,
 * written to rows row0 + i; gen_row0 is the offset in the global synthetic corpus) */
int rag_corpus_fill_synthetic(rag_corpus_t* c, uint64_t seed, int64_t gen_row0, int64_t row0, int64_t nrows);
int rag_corpus_device_ptr(const rag_corpus_t* c, void** rows_dev);   /* single-shard corpora */

/* ---- dense similarity + exact top-k ---------------------------------------
 * replaces collection.query(query_embeddings=[vec], n_results=k, where=...)
 * at src/rag/retriever.py:215-220 and :380-385 (cosine space,
 * src/processing/create_chromadb_index.py:100-106).
 *   q            B x dim fp32, row-major (already L2-normalised by the caller)
 *   allow_bitmap NULL or ceil(count/8) bytes, bit r set = row r passes `where`
 *   out_rows     B x k row indices (global rows of a sharded corpus), -1 padded
 *   out_scores   B x k canonical fp64 <q,x>  (distance = 1 - score)
 *   out_counts   B     number of valid results (< k if the filter leaves fewer)
 * Tombstoned rows never appear. */
int rag_dense_topk(rag_corpus_t* c, const float* q, int B, int k, const uint8_t* allow_bitmap,
                   int32_t* out_rows, double* out_scores, int32_t* out_counts);
/* same, the filter is a compiled `where` predicate over the coded columns (csrc/rowfilter.cu): postfix program
 *   0 TRUE | 1 FALSE | 2 EQ col code | 3 NE col code | 4 IN col nbits nwords w0.. | 5 NIN col nbits nwords w0.. |
 *   6 AND n | 7 OR n
 * (the `where` shapes of src/rag/pipeline.py:35-71; a missing key fails EQ / IN and passes NE / NIN).  The
 * predicate is evaluated ON THE DEVICE into a row bitmap that is cached per (program, corpus version): the steady
 * state uploads nothing but the program text.  where_prog == NULL or n_words == 0: no filter. */
int rag_dense_topk_where(rag_corpus_t* c, const float* q, int B, int k, const int32_t* where_prog, int n_words,
                         int32_t* out_rows, double* out_scores, int32_t* out_counts);
/* every pointer is a DEVICE pointer (inputs resident in HBM); single-shard corpora.  STREAM-ORDERED: the call
 * queues the filter, the exact refine AND a device-driven exact fallback pass for queries whose margin check fails,
 * and returns; results are ready when the stream (rag_set_stream) reaches that point.  A batch in which more
 * than 8 queries need the fallback at once marks the surplus with out_counts = -1 (rag_dense_topk redoes them). */
int rag_dense_topk_dev(rag_corpus_t* c, const float* q_dev, int B, int k, const uint8_t* allow_bitmap_dev,
                       int32_t* out_rows_dev, double* out_scores_dev, int32_t* out_counts_dev);
/* multi-GPU exchange step (one process per GPU): after an all-gather of the G
 * ranks' (score, global id) lists, keep the global top-k per query, order
 * (score desc, id asc).  Device pointers: rank g's B x k block of scores / ids
 * starts at element g * rank_stride (0 = B*k, i.e. dense G x B x k arrays), so one
 * packed all-gather buffer can carry both.  STREAM-ORDERED. */
int rag_merge_topk_dev(const double* scores_dev, const int64_t* ids_dev, int G, int B, int k, int64_t rank_stride,
                       double* out_scores_dev, int64_t* out_ids_dev, int32_t* out_counts_dev);

/* The same exchange step WITHOUT a collective library: every rank stores its (score, global id) block straight
 * into every peer's gather buffer over NVLink peer memory (cudaIpc-mapped), raises an epoch flag there, and the
 * merge kernel of each rank waits for the G flags before it merges (csrc/exchange.cu).  Set-up, once:
 *   rag_exchange_create   allocates this rank's buffer (slot_bytes >= B*k*16 of the largest call) and returns
 *                         its RAG_IPC_HANDLE_BYTES-byte handle; exchange the handles between the processes by
 *                         any means (e.g. torch.distributed.all_gather_object);
 *   rag_exchange_connect  handles = world x RAG_IPC_HANDLE_BYTES bytes, rank-major (this rank's own entry is
 *                         ignored).
 * Every rank's list must be what rag_dense_topk(_dev) returns: sorted by (score desc, id asc), padding (-1) at the end,
 * ids of different ranks disjoint — the merge RANKS the entries by binary search in the other lists instead of sorting.
 * rag_exchange_merge_*_dev are STREAM-ORDERED (two launches, no host synchronisation); all ranks must call one
 * of them once per step with the same B and k, always on the same stream.  A rank that never arrives does not
 * hang the others: after "exchange_timeout_ms" their queries of that step carry out_counts = -2 and
 * rag_exchange_status reports it. */
#define RAG_IPC_HANDLE_BYTES 64
typedef struct rag_exchange rag_exchange_t;
int rag_exchange_create(rag_exchange_t** out, int world, int rank, size_t slot_bytes, void* handle_out);
int rag_exchange_connect(rag_exchange_t* ex, const void* handles);
int rag_exchange_destroy(rag_exchange_t* ex);
int rag_exchange_merge_topk_dev(rag_exchange_t* ex, const double* my_scores_dev, const int64_t* my_ids_dev, int B,
                                int k, double* out_scores_dev, int64_t* out_ids_dev, int32_t* out_counts_dev);
/* same, the rank hands over what rag_dense_topk_dev wrote: LOCAL int32 rows — the push kernel turns them into
 * global ids (row_lo + row, -1 stays padding) on the way out — and (nullable) its counts: a query some rank
 * left unresolved (count -1) comes out with out_counts = -1 on EVERY rank, so that all ranks redo it together */
int rag_exchange_merge_rows_dev(rag_exchange_t* ex, const double* my_scores_dev, const int32_t* my_rows_dev,
                                int64_t row_lo, const int32_t* my_counts_dev, int B, int k, double* out_scores_dev,
                                int64_t* out_ids_dev, int32_t* out_counts_dev);
int rag_exchange_status(rag_exchange_t* ex, int* timed_out);

/* ---- BM25 keyword scoring over CSR postings -------------------------------
 * replaces rank_bm25.BM25Okapi(corpus_tokens) / .get_scores(tokens)
 * (src/rag/bm25_index.py:126,153,236,265) and the select loop of
 * ChunkBM25Index.search (src/rag/bm25_index.py:267-279).
 *   term_ptr  n_terms+1 offsets into the postings, postings sorted by row
 *   idf       per-term weight AFTER the negative-idf epsilon floor
 * fp64 arithmetic in numpy's evaluation order; scores bit-identical. */
int rag_bm25_create(rag_bm25_t** out, int64_t n_docs, int64_t n_terms, int64_t nnz, const int64_t* term_ptr,
                    const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, const double* idf,
                    double avgdl, double k1, double b);
/* the same index with its rows (documents) spread over the first n_shards shard slots like a sharded corpus:
 * every shard holds the postings of its rows; idf and avgdl are the GLOBAL statistics passed in */
int rag_bm25_create_sharded(rag_bm25_t** out, int n_shards, int64_t n_docs, int64_t n_terms, int64_t nnz,
                            const int64_t* term_ptr, const int32_t* post_row, const int32_t* post_tf,
                            const int32_t* doc_len, const double* idf, double avgdl, double k1, double b);
int rag_bm25_destroy(rag_bm25_t* ix);
/* measurement helper: the bytes the filter pass streams for each query = sum over its scoring tokens of the term's
 * list in the index format that serves it (4 per posting of the packed stream, 2 per row of a dense column; 12 per
 * posting when only the exact path exists), and the postings themselves (sum of df). Either output may be NULL. */
int rag_bm25_query_bytes(const rag_bm25_t* ix, const int32_t* q_terms, const int32_t* q_ptr, int Q, int64_t* out_bytes,
                         int64_t* out_postings);
/* Q queries; q_terms are the concatenated term ids (in token order, repeats
 * kept, -1 = token outside the vocabulary), q_ptr has Q+1 offsets.
 * allow_bitmap: NULL or one bitmap shared by all queries (doc_filter).
 * Results: score > 0 only, score desc, ties -> lowest row; -1 padded. */
int rag_bm25_search(rag_bm25_t* ix, const int32_t* q_terms, const int32_t* q_ptr, int Q, int k,
                    const uint8_t* allow_bitmap, int32_t* out_rows, double* out_scores, int32_t* out_counts);
/* full score vector of one query (BM25Okapi.get_scores parity), n_docs doubles */
int rag_bm25_scores(rag_bm25_t* ix, const int32_t* q_terms, int n_q_terms, double* out_scores);
/* bytes per posting the search kernels stream (4: packed filter path; 12: exact fp64 path only) */
int rag_bm25_info(const rag_bm25_t* ix, int* bytes_per_posting, int64_t* index_bytes);

/* ---- weighted Reciprocal Rank Fusion --------------------------------------
 * replaces reciprocal_rank_fusion (src/rag/retriever.py:66-90) + the stable
 * descending sort and cut of the fusion tail (:464-467) for Q questions at once.
 *   ids      Q x R x L integer ids, negative = padding (skipped, rank not advanced)
 *   weights  Q x R
 * Output per question: distinct ids by (fused score desc, first-seen order),
 * at most `top`; -1 padded.  R * L <= 8192. */
int rag_rrf_fuse(const int32_t* ids, const double* weights, int Q, int R, int L, int rrf_k, int top,
                 int32_t* out_ids, double* out_scores, int32_t* out_counts);

/* ---- rerank select: the output step of the cross-encoder reranker -----------
 * replaces the tail of CrossEncoderReranker.rerank (src/rag/reranker.py:172-211; called at
 * src/rag/pipeline.py:250-256 right after retrieve_candidates) for Q questions at once. The cross-encoder itself is
 * model inference and stays outside. scores: Q x L fp32 model scores in candidate order (what CrossEncoder.predict
 * returns), boosts: Q x L fp64 topic boosts (NULL: none), lens: candidates per question (NULL: L each).
 * final = (double)score + boost; stable descending order; first top_k; entries below min_score dropped, but at
 * least 3 are returned when 3 candidates exist (so top_k >= 3 is required). out_idx: Q x top_k positions in the
 * candidate list (-1 padded), out_scores: their final scores, out_counts: entries per question. L <= 1024. */
int rag_rerank_select(const float* scores, const double* boosts, const int32_t* lens, int Q, int L, int top_k,
                      double min_score, int32_t* out_idx, double* out_scores, int32_t* out_counts);

/* ---- host-side index construction (multi-threaded C++, no GPU needed) ----------------------------------
 * replaces the per-document dict building of rank_bm25.BM25Okapi._initialize (rank-bm25 0.2.2; reached from
 * ChunkBM25Index.build_from_collection, src/rag/bm25_index.py:236, and SummaryBM25Index.build, :126):
 * documents as concatenated term ids (doc_ptr: n_docs+1 offsets) -> CSR postings (term_ptr n_terms+1,
 * post_row ascending per term, post_tf).  post_row / post_tf hold `capacity` entries (doc_ptr[n_docs] always
 * suffices); *nnz_out = entries written.  n_threads <= 0: all host cores. */
int rag_csr_build(int64_t n_docs, const int64_t* doc_ptr, const int32_t* tokens, int64_t n_terms, int64_t* term_ptr,
                  int32_t* post_row, int32_t* post_tf, int64_t capacity, int64_t* nnz_out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* B200RAG_H */
