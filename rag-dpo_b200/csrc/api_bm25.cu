// api_bm25.cu — the BM25 entry points of the C ABI (include/b200rag.h): index construction on the device (exact
// fp64 impacts + the packed filter index of the fast path), search (single GPU and sharded over the GPUs of one
// process) and the full score vector.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "api_common.h"

using namespace b200rag;
namespace b200rag { extern int g_bm25_tile_chunks; extern int g_bm25_tma; extern int g_bm25_by_block; extern int g_bm25_acc16; extern int g_bm25_inline_resolve; }      // bm25.cu: options "bm25_tile", "bm25_tma"

namespace {

int g_bm25_fast = 1;          // option "bm25_fast": 0 = always the exact fp64 range path
int g_bm25_rows_max = kBm25MaxListedRows;   // option "bm25_rows_max": largest row filter served by the listed-rows kernel
int g_bm25_dense_div = 8;     // option "bm25_dense_div": terms with df >= n_docs / div get a dense column (0 = none)

struct Bm25Shard {
    Ctx cx;
    Bm25Device d{};
    int64_t index_bytes = 0;
    std::vector<int64_t> h_term_ptr;
    std::vector<double> h_idf;
    std::vector<uint8_t> h_cls;        // term class of the filter index (kBmLow / kBmMid / kBmDense / kBmSkip); empty: no filter index
    // tables of the fast path (device)
    DevBuf tabled_terms, dense_terms;
    int n_dense = 0, n_tabled = 0;
    // scratch
    DevBuf inbuf, outbuf, cand, scratch, index, ranges;
    // views into inbuf (terms | q_ptr | allow bitmap | listed rows: ONE upload) and outbuf (scores | rows | counts:
    // ONE download), set by shard_search_launch
    int32_t *d_terms = nullptr, *d_qptr = nullptr, *d_listed = nullptr, *d_rows = nullptr, *d_counts = nullptr;
    uint8_t* d_allow = nullptr;
    double* d_scores = nullptr;
};

void shard_free(Bm25Shard* s) {
    if (!s) return;
    if (s->cx.ready) {
        cudaSetDevice(s->cx.device);
        cudaStreamSynchronize(s->cx.stream());
    }
    cudaFree(s->d.term_ptr); cudaFree(s->d.post_row); cudaFree(s->d.post_impact); cudaFree(s->d.idf); cudaFree(s->d.score);
    cudaFree(s->d.post_pack); cudaFree(s->d.term_info); cudaFree(s->d.rng_off); cudaFree(s->d.dense_col);
    for (DevBuf* b : {&s->tabled_terms, &s->dense_terms, &s->inbuf, &s->outbuf, &s->cand, &s->scratch, &s->index, &s->ranges})
        b->release();
    {
        std::lock_guard<std::mutex> tl(R.tmu);
        if (g_last_ctx == &s->cx) g_last_ctx = nullptr;
    }
    s->cx.destroy();
    cudaGetLastError();
    delete s;
}

// builds one shard on shard slot `slot` from a CSR whose rows are LOCAL to the shard
int shard_build(Bm25Shard** out, int slot, int64_t n_docs, int64_t n_terms, int64_t nnz, const int64_t* term_ptr,
                const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, const double* idf, double avgdl,
                double k1, double b) {
    Bm25Shard* s = new Bm25Shard();
    int rc = s->cx.init(R.slots[slot]);
    if (rc != RAG_OK) { delete s; return rc; }
    cudaStream_t st = s->cx.stream();
    Bm25Device& d = s->d;
    d.n_docs = n_docs;
    d.n_terms = n_terms;
    d.nnz = nnz;
    d.n_blocks = (int)((n_docs + kBm25Block - 1) / kBm25Block);
    if (n_docs == 0) nnz = 0;
    s->h_term_ptr.assign(term_ptr, term_ptr + n_terms + 1);
    s->h_idf.assign(idf, idf + n_terms);
    int32_t *d_tf = nullptr, *d_dl = nullptr;
    unsigned long long* d_max = nullptr;
    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    cudaError_t e = cudaSuccess;
    int64_t bytes = 0;
    auto A = [&](void** p, size_t n) {
        if (e == cudaSuccess) { e = cudaMalloc(p, n); bytes += (int64_t)n; }
    };
    auto C = [&](void* dst, const void* src, size_t n) {
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
    };
    auto bail = [&](const char* what) {
        const cudaError_t err = e;
        cudaFree(d_tf); cudaFree(d_dl); cudaFree(d_max);
        shard_free(s);
        cudaGetLastError();
        return fail(err == cudaErrorMemoryAllocation ? RAG_ENOMEM : RAG_ECUDA, "bm25 build (%s): %s", what, cudaGetErrorString(err));
    };
    A((void**)&d.term_ptr, (size_t)(n_terms + 1) * 8);
    A((void**)&d.post_row, nz * 4);
    A((void**)&d.post_impact, nz * 8);
    A((void**)&d.idf, (size_t)std::max<int64_t>(n_terms, 1) * 8);
    A((void**)&d.score, (size_t)std::max<int64_t>(n_docs, 1) * 8);
    A((void**)&d_tf, nz * 4);
    A((void**)&d_dl, (size_t)std::max<int64_t>(n_docs, 1) * 4);
    A((void**)&d_max, 8);
    C(d.term_ptr, term_ptr, (size_t)(n_terms + 1) * 8);
    C(d.post_row, post_row, (size_t)nnz * 4);
    C(d_tf, post_tf, (size_t)nnz * 4);
    C(d_dl, doc_len, (size_t)n_docs * 4);
    C(d.idf, idf, (size_t)n_terms * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(d.score, 0, (size_t)std::max<int64_t>(n_docs, 1) * 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_max, 0, 8, st);
    if (e == cudaSuccess && nnz > 0) {
        e = bm25_impact_launch(d.post_row, d_tf, d_dl, nnz, avgdl, k1, b, d.post_impact, st);
        ++R.n_launch;
        if (e == cudaSuccess) e = bm25_cmax_launch(d, d_max, st);
        ++R.n_launch;
    }
    unsigned long long h_max = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_max, d_max, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail("postings");
    cudaFree(d_tf); d_tf = nullptr;
    cudaFree(d_dl); d_dl = nullptr;
    cudaFree(d_max); d_max = nullptr;
    // ---- the packed filter index of the fast path.  Valid only when no idf is negative (the negative-idf
    // epsilon floor of rank-bm25 turns negative only in degenerate corpora; those take the exact path).
    double max_impact = 0.0, idf_max = 0.0;
    memcpy(&max_impact, &h_max, 8);
    bool any_negative = false;
    for (int64_t t = 0; t < n_terms; ++t) {
        if (idf[t] < 0.0) any_negative = true;
        if (idf[t] > idf_max) idf_max = idf[t];
    }
    const double cbound = idf_max * max_impact;          // >= idf[t] * impact[p] for every posting (rn is monotone)
    d.fast_ok = 0;
    // (the filter kernel addresses the packed stream with 32-bit offsets)
    if (nnz > 0 && nnz < 0xFFFF0000LL && d.n_blocks >= 1 && cbound > 0.0 && std::isfinite(cbound)) {
        // 2^18 - 144: q = ceil(c / unit) + 1 stays inside 18 bits and ceil(q / 4) inside 16
        const double unit = cbound / 262000.0;
        // term classes by document frequency.  DENSE: df >= n_docs / g_bm25_dense_div, the most frequent first, while
        // the columns (2 bytes x rows each) fit the budget of one more copy of the packed stream.  Tabled (range
        // table row of (n_blocks + 1) x 4 bytes): df >= mid_min, raised until the tables fit a quarter of the stream.
        const int64_t col_rows = (int64_t)d.n_blocks * kBm25Block;
        const int64_t ddiv = std::max(g_bm25_dense_div, 1);
        const int64_t dense_min = std::max<int64_t>((n_docs + ddiv - 1) / ddiv, 64);
        std::vector<std::pair<int64_t, int32_t>> dense_cand;
        std::vector<int64_t> dfs;
        dfs.reserve((size_t)n_terms);
        for (int64_t t = 0; t < n_terms; ++t) {
            const int64_t df = term_ptr[t + 1] - term_ptr[t];
            if (df > 0 && idf[t] != 0.0) {
                dfs.push_back(df);
                if (g_bm25_dense_div > 0 && df >= dense_min) dense_cand.emplace_back(-df, (int32_t)t);
            }
        }
        std::sort(dense_cand.begin(), dense_cand.end());
        const int64_t col_budget = std::max<int64_t>(nnz * 4, 64LL << 20);
        size_t n_dense = 0;
        while (n_dense < dense_cand.size() && n_dense < 4096 && (int64_t)(n_dense + 1) * col_rows * 2 <= col_budget) ++n_dense;
        dense_cand.resize(n_dense);
        int64_t mid_min = std::min<int64_t>(std::max<int64_t>(d.n_blocks, 64), 256);
        {
            std::sort(dfs.begin(), dfs.end(), std::greater<int64_t>());
            const int64_t tab_budget = std::max<int64_t>(nnz, 16LL << 20);       // bytes: a quarter of the packed stream
            const int64_t max_tabled = tab_budget / (4 * ((int64_t)d.n_blocks + 1));
            if ((int64_t)dfs.size() > max_tabled && max_tabled >= 0) {
                const int64_t cut = max_tabled > 0 ? dfs[(size_t)max_tabled - 1] : dfs[0] + 1;
                mid_min = std::max(mid_min, cut + 1);
            }
        }
        std::vector<int2> info((size_t)std::max<int64_t>(n_terms, 1));
        std::vector<int32_t> tabled, dense;
        std::vector<int32_t> col_of((size_t)std::max<int64_t>(n_terms, 1), -1);
        for (size_t i = 0; i < dense_cand.size(); ++i) {
            col_of[(size_t)dense_cand[i].second] = (int32_t)i;
            dense.push_back(dense_cand[i].second);
        }
        for (int64_t t = 0; t < n_terms; ++t) {
            const int64_t df = term_ptr[t + 1] - term_ptr[t];
            int cls = kBmLow, slot_t = 0, col = 0;
            if (df == 0 || idf[t] == 0.0) {
                cls = kBmSkip;
            } else if (col_of[(size_t)t] >= 0 || df >= mid_min) {
                cls = col_of[(size_t)t] >= 0 ? kBmDense : kBmMid;
                col = std::max(col_of[(size_t)t], 0);
                slot_t = (int)tabled.size();
                tabled.push_back((int32_t)t);
            }
            info[(size_t)t] = make_int2((int)(((unsigned)cls << 30) | (unsigned)slot_t), col);
        }
        const size_t rng_n = (size_t)std::max<size_t>(tabled.size(), 1) * (d.n_blocks + 1);
        const size_t col_n = (size_t)std::max<size_t>(dense.size(), 1) * (size_t)col_rows;
        A((void**)&d.post_pack, (nz + 4) * 4);
        A((void**)&d.term_info, info.size() * sizeof(int2));
        A((void**)&d.rng_off, rng_n * 4);
        A((void**)&d.dense_col, col_n * 2);
        int rc2 = RAG_OK;
        if (e == cudaSuccess) rc2 = s->tabled_terms.ensure(std::max<size_t>(tabled.size(), 1) * 4);
        if (e == cudaSuccess && rc2 == RAG_OK) rc2 = s->dense_terms.ensure(std::max<size_t>(dense.size(), 1) * 4);
        if (rc2 != RAG_OK) { shard_free(s); return rc2; }
        bytes += (int64_t)(s->tabled_terms.bytes + s->dense_terms.bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(d.post_pack + nz, 0, 16, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d.dense_col, 0, col_n * 2, st);
        C(d.term_info, info.data(), info.size() * sizeof(int2));
        C(s->tabled_terms.p, tabled.data(), tabled.size() * 4);
        C(s->dense_terms.p, dense.data(), dense.size() * 4);
        if (e == cudaSuccess) {
            e = bm25_index_build_launch(d, unit, s->tabled_terms.as<int32_t>(), (int)tabled.size(), s->dense_terms.as<int32_t>(),
                                        (int)dense.size(), st);
            R.n_launch += 3;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return bail("filter index");
        s->n_dense = (int)dense.size();
        d.n_dense = (int)dense.size();
        s->n_tabled = (int)tabled.size();
        s->h_cls.resize((size_t)n_terms);
        for (int64_t t = 0; t < n_terms; ++t) s->h_cls[(size_t)t] = (uint8_t)((unsigned)info[(size_t)t].x >> 30);
        d.fast_ok = any_negative ? 0 : 1;
    }
    s->index_bytes = bytes;
    *out = s;
    return RAG_OK;
}

}  // namespace

struct rag_bm25 {
    int n_shards = 1;
    int64_t n_docs = 0, n_terms = 0;
    std::vector<Bm25Shard*> sh;
    std::recursive_mutex mu;
    DevBuf g_scores, g_ids, m_scores, m_ids, m_counts;     // sharded search: gather buffers on the primary shard
};

namespace {

struct IndexLock {
    rag_bm25* ix;
    explicit IndexLock(const rag_bm25* i) : ix(const_cast<rag_bm25*>(i)) {
        ix->mu.lock();
        for (auto* s : ix->sh) s->cx.mu.lock();
    }
    ~IndexLock() {
        for (auto it = ix->sh.rbegin(); it != ix->sh.rend(); ++it) (*it)->cx.mu.unlock();
        ix->mu.unlock();
    }
};

void index_free(rag_bm25* ix) {
    for (auto* s : ix->sh) shard_free(s);
    for (DevBuf* b : {&ix->g_scores, &ix->g_ids, &ix->m_scores, &ix->m_ids, &ix->m_counts}) b->release();
    delete ix;
}

int create_check(rag_bm25_t** out, int64_t n_docs, int64_t n_terms, int64_t nnz, const int64_t* term_ptr,
                 const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, const double* idf) {
    RAG_TRY(require_init());
    if (!out || !term_ptr || !doc_len || !idf || (nnz > 0 && (!post_row || !post_tf))) return fail(RAG_EINVAL, "NULL argument");
    if (n_docs <= 0 || n_terms < 0 || nnz < 0 || n_docs > 0x7FFFFFF0LL) return fail(RAG_EINVAL, "bad sizes");
    if (term_ptr[0] != 0 || term_ptr[n_terms] != nnz) return fail(RAG_EINVAL, "term_ptr does not span the postings");
    return RAG_OK;
}

// queues one shard's search: uploads the queries, runs the fast path (or the listed-rows / exact range path) and
// leaves rows (LOCAL) / scores / counts in the shard's result buffers.  rows_list != nullptr: the allowed rows.
struct SearchPlan {
    bool fast = false, listed = false;
    int kp = 0, n_lists = 0;
    const uint8_t* allow_dev = nullptr;
};

int shard_search_launch(Bm25Shard& s, const int32_t* q_terms, const int32_t* q_ptr, int Q, int k, int max_tokens,
                        const uint8_t* allow_local, const int32_t* listed_rows, int n_listed, SearchPlan& plan) {
    RAG_TRY(s.cx.use());
    cudaStream_t st = s.cx.stream();
    const Bm25Device& d = s.d;
    const int n_tok = q_ptr[Q];
    const size_t tb = (size_t)std::max(n_tok, 1) * 4, pb = (size_t)(Q + 1) * 4;
    const size_t ab = allow_local ? (size_t)((d.n_docs + 7) / 8) : 0;
    const size_t lb = listed_rows ? (size_t)std::max(n_listed, 1) * 4 : 0;
    const size_t rb = (size_t)Q * k * 4, sb = (size_t)Q * k * 8, cb = (size_t)Q * 4;
    plan.kp = std::max(16, next_pow2(k));
    plan.n_lists = bm25_range_lists(d.n_docs);
    auto up16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
    const size_t o_qptr = up16(tb), o_allow = o_qptr + up16(pb), o_listed = o_allow + up16(ab ? ab + 16 : 0);
    const size_t in_bytes = o_listed + up16(lb);
    RAG_TRY(s.inbuf.ensure(in_bytes));
    RAG_TRY(s.outbuf.ensure(sb + rb + cb));
    s.d_terms = s.inbuf.as<int32_t>();
    s.d_qptr = reinterpret_cast<int32_t*>(s.inbuf.as<uint8_t>() + o_qptr);
    s.d_allow = s.inbuf.as<uint8_t>() + o_allow;
    s.d_listed = reinterpret_cast<int32_t*>(s.inbuf.as<uint8_t>() + o_listed);
    s.d_scores = s.outbuf.as<double>();
    s.d_rows = reinterpret_cast<int32_t*>(s.outbuf.as<uint8_t>() + sb);
    s.d_counts = reinterpret_cast<int32_t*>(s.outbuf.as<uint8_t>() + sb + rb);
    RAG_TRY(s.cx.ensure_pinned(std::max(in_bytes, sb + rb + cb)));
    uint8_t* pin = reinterpret_cast<uint8_t*>(s.cx.pinned);
    if (n_tok > 0) memcpy(pin, q_terms, (size_t)n_tok * 4);
    memcpy(pin + o_qptr, q_ptr, pb);
    if (ab) memcpy(pin + o_allow, allow_local, ab);
    if (lb && n_listed > 0) memcpy(pin + o_listed, listed_rows, (size_t)n_listed * 4);
    CU_TRY(cudaMemcpyAsync(s.inbuf.p, pin, in_bytes, cudaMemcpyHostToDevice, st));
    plan.allow_dev = ab ? s.d_allow : nullptr;
    s.cx.clear_timing();
    set_last_ctx(&s.cx);
    s.cx.rec(0);
    if (d.n_docs == 0) {                       // an empty shard of a small sharded index
        CU_TRY(cudaMemsetAsync(s.d_rows, 0xFF, rb, st));
        CU_TRY(cudaMemsetAsync(s.d_scores, 0, sb, st));
        CU_TRY(cudaMemsetAsync(s.d_counts, 0, cb, st));
        plan.fast = plan.listed = false;
        s.cx.rec(1);
        return RAG_OK;
    }
    plan.listed = listed_rows != nullptr;
    plan.fast = !plan.listed && g_bm25_fast && bm25_fast_supported(d, k, max_tokens);
    if (plan.listed) {
        // selective row filter: exact scores of the listed rows only — no posting stream at all
        CU_TRY(bm25_rows_launch(d, s.d_terms, s.d_qptr, Q, s.d_listed, n_listed, k,
                                s.d_rows, s.d_scores, s.d_counts, st));
        ++R.n_launch;
    } else if (plan.fast) {
        const int chunk = bm25_fast_chunk(d, max_tokens);  // scratch budget / gridDim.y limit
        for (int q0 = 0; q0 < Q; q0 += chunk) {
            const int nq = std::min(chunk, Q - q0);
            RAG_TRY(s.scratch.ensure(bm25_fast_scratch_bytes(d, k, nq, max_tokens)));
            CU_TRY(bm25_fast_launch(d, s.d_terms, s.d_qptr, q0, nq, max_tokens, plan.allow_dev, k, s.scratch.p,
                                    s.d_rows, s.d_scores, s.d_counts, st));
            R.n_launch += 3;
        }
    } else {
        RAG_TRY(s.cand.ensure((size_t)Q * plan.n_lists * plan.kp * bm25_key_bytes()));
        CU_TRY(bm25_range_launch(d, s.d_terms, s.d_qptr, nullptr, Q, plan.allow_dev, plan.kp, k,
                                 s.cand.p, s.d_rows, s.d_scores, s.d_counts, st));
        R.n_launch += 2;
    }
    s.cx.rec(1);
    return RAG_OK;
}

// after the shard's stream was synchronised and h_counts holds its counts: queries the fast path flagged
// (count = -1: mass ties at the bound) are redone on the exact range path, in place.  Returns the number redone.
int shard_search_redo(Bm25Shard& s, int Q, int k, const SearchPlan& plan, int32_t* h_counts, int* n_redone) {
    *n_redone = 0;
    if (!plan.fast) return RAG_OK;
    std::vector<int32_t> redo;
    for (int q = 0; q < Q; ++q)
        if (h_counts[q] < 0) redo.push_back(q);
    if (redo.empty()) return RAG_OK;
    if (getenv("B200RAG_DEBUG")) {
        fprintf(stderr, "[b200rag] bm25: %zu of %d queries redone on the exact range path:", redo.size(), Q);
        for (size_t i = 0; i < redo.size() && i < 16; ++i) fprintf(stderr, " %d", redo[i]);
        fprintf(stderr, "\n");
    }
    RAG_TRY(s.cx.use());
    cudaStream_t st = s.cx.stream();
    ++R.n_fallback;
    const int nr = (int)redo.size();
    DevBuf r_rows, r_scores, r_counts;
    auto done = [&](int rc) {
        r_rows.release(); r_scores.release(); r_counts.release();
        return rc;
    };
    int rc = s.index.ensure((size_t)nr * 4);
    if (rc == RAG_OK) rc = s.cand.ensure((size_t)nr * plan.n_lists * plan.kp * bm25_key_bytes());
    if (rc == RAG_OK) rc = r_rows.ensure((size_t)nr * k * 4);
    if (rc == RAG_OK) rc = r_scores.ensure((size_t)nr * k * 8);
    if (rc == RAG_OK) rc = r_counts.ensure((size_t)nr * 4);
    if (rc != RAG_OK) return done(rc);
    cudaError_t e = cudaMemcpyAsync(s.index.p, redo.data(), (size_t)nr * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = bm25_range_launch(s.d, s.d_terms, s.d_qptr, s.index.as<int32_t>(), nr, plan.allow_dev, plan.kp, k,
                              s.cand.p, r_rows.as<int32_t>(), r_scores.as<double>(), r_counts.as<int32_t>(), st);
    R.n_launch += 2;
    // scatter the redone results into the shard's result buffers
    for (int i = 0; i < nr && e == cudaSuccess; ++i) {
        const int q = redo[i];
        e = cudaMemcpyAsync(s.d_rows + (size_t)q * k, r_rows.as<int32_t>() + (size_t)i * k, (size_t)k * 4,
                            cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(s.d_scores + (size_t)q * k, r_scores.as<double>() + (size_t)i * k, (size_t)k * 8,
                                cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(s.d_counts + q, r_counts.as<int32_t>() + i, 4, cudaMemcpyDeviceToDevice, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(fail(RAG_ECUDA, "bm25 redo: %s", cudaGetErrorString(e)));
    *n_redone = nr;
    return done(RAG_OK);
}

__global__ void bm25_publish_kernel(const int32_t* __restrict__ rows, const double* __restrict__ scores, int64_t n, int shard,
                                    int n_shards, int64_t* __restrict__ g_ids, double* __restrict__ g_scores) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t r = rows[i];
        g_ids[i] = r < 0 ? (int64_t)-1 : shard_global_row(r, shard, n_shards, RAG_SHARD_BLOCK);
        g_scores[i] = scores[i];
    }
}

// launches the per-token accumulation of one query on one shard; fills `ranges` with the [lo,hi) posting ranges of
// its distinct scoring tokens
int accumulate_query(Bm25Shard& s, const int32_t* terms, int nt, std::vector<int64_t>& ranges, int64_t* total) {
    ranges.clear();
    *total = 0;
    std::vector<int32_t> seen;
    cudaStream_t st = s.cx.stream();
    for (int i = 0; i < nt; ++i) {
        const int32_t t = terms[i];
        if (t < 0 || t >= s.d.n_terms) continue;          // (idf.get(q) or 0) == 0
        const double w = s.h_idf[t];
        const int64_t lo = s.h_term_ptr[t], hi = s.h_term_ptr[t + 1];
        if (w == 0.0 || hi <= lo) continue;
        CU_TRY(bm25_accumulate_launch(s.d, lo, hi, w, st));
        ++R.n_launch;
        if (std::find(seen.begin(), seen.end(), t) == seen.end()) {
            seen.push_back(t);
            ranges.push_back(lo);
            ranges.push_back(hi);
            *total += hi - lo;
        }
    }
    return RAG_OK;
}

}  // namespace

int bm25_set_option(const char* key, int64_t value) {
    if (!strcmp(key, "bm25_fast")) g_bm25_fast = value != 0;
    else if (!strcmp(key, "bm25_tile")) b200rag::g_bm25_tile_chunks = (int)value;
    else if (!strcmp(key, "bm25_tma")) b200rag::g_bm25_tma = (int)value;
    else if (!strcmp(key, "bm25_by_block")) b200rag::g_bm25_by_block = (int)value;
    else if (!strcmp(key, "bm25_acc16")) b200rag::g_bm25_acc16 = (int)value;
    else if (!strcmp(key, "bm25_inline_resolve")) b200rag::g_bm25_inline_resolve = (int)value;
    else if (!strcmp(key, "bm25_dense_div")) g_bm25_dense_div = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20));
    else if (!strcmp(key, "bm25_rows_max")) g_bm25_rows_max = (int)std::max<int64_t>(0, std::min<int64_t>(value, kBm25MaxListedRows));
    else return RAG_EINVAL;
    return RAG_OK;
}

extern "C" {

int rag_bm25_create(rag_bm25_t** out, int64_t n_docs, int64_t n_terms, int64_t nnz, const int64_t* term_ptr,
                    const int32_t* post_row, const int32_t* post_tf, const int32_t* doc_len, const double* idf,
                    double avgdl, double k1, double b) {
    RAG_TRY(create_check(out, n_docs, n_terms, nnz, term_ptr, post_row, post_tf, doc_len, idf));
    rag_bm25* ix = new rag_bm25();
    ix->n_shards = 1;
    ix->n_docs = n_docs;
    ix->n_terms = n_terms;
    Bm25Shard* s = nullptr;
    int rc = shard_build(&s, 0, n_docs, n_terms, nnz, term_ptr, post_row, post_tf, doc_len, idf, avgdl, k1, b);
    if (rc != RAG_OK) { delete ix; return rc; }
    ix->sh.push_back(s);
    *out = ix;
    return RAG_OK;
}

int rag_bm25_create_sharded(rag_bm25_t** out, int n_shards, int64_t n_docs, int64_t n_terms, int64_t nnz,
                            const int64_t* term_ptr, const int32_t* post_row, const int32_t* post_tf,
                            const int32_t* doc_len, const double* idf, double avgdl, double k1, double b) {
    RAG_TRY(create_check(out, n_docs, n_terms, nnz, term_ptr, post_row, post_tf, doc_len, idf));
    if (n_shards < 1 || n_shards > (int)R.slots.size())
        return fail(RAG_EINVAL, "n_shards=%d but %zu shard slot(s) are initialised (rag_init_devices)", n_shards, R.slots.size());
    if (n_shards == 1) return rag_bm25_create(out, n_docs, n_terms, nnz, term_ptr, post_row, post_tf, doc_len, idf, avgdl, k1, b);
    const int G = n_shards;
    // split the CSR by the block-cyclic row layout: every shard keeps the postings of its rows (local row numbers,
    // still ascending per term); idf and avgdl stay the global statistics
    std::vector<std::vector<int64_t>> tptr(G, std::vector<int64_t>((size_t)n_terms + 1, 0));
    for (int64_t t = 0; t < n_terms; ++t)
        for (int64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
            if (post_row[p] < 0 || post_row[p] >= n_docs) return fail(RAG_EINVAL, "posting row out of range");
            ++tptr[shard_of_row(G, post_row[p])][(size_t)t + 1];
        }
    std::vector<std::vector<int32_t>> rows(G), tfs(G), dls(G);
    for (int s = 0; s < G; ++s) {
        for (int64_t t = 0; t < n_terms; ++t) tptr[s][(size_t)t + 1] += tptr[s][(size_t)t];
        rows[s].resize((size_t)std::max<int64_t>(tptr[s][(size_t)n_terms], 1));
        tfs[s].resize(rows[s].size());
        dls[s].assign((size_t)std::max<int64_t>(shard_row_count(G, s, n_docs), 1), 0);
    }
    {
        std::vector<std::vector<int64_t>> cur(G);
        for (int s = 0; s < G; ++s) cur[s].assign(tptr[s].begin(), tptr[s].end() - 1);
        for (int64_t t = 0; t < n_terms; ++t)
            for (int64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
                const int s = shard_of_row(G, post_row[p]);
                const int64_t pos = cur[s][(size_t)t]++;
                rows[s][(size_t)pos] = (int32_t)local_of_row(G, post_row[p]);
                tfs[s][(size_t)pos] = post_tf[p];
            }
    }
    for (int64_t r = 0; r < n_docs; ++r) dls[shard_of_row(G, r)][(size_t)local_of_row(G, r)] = doc_len[r];
    rag_bm25* ix = new rag_bm25();
    ix->n_shards = G;
    ix->n_docs = n_docs;
    ix->n_terms = n_terms;
    for (int s = 0; s < G; ++s) {
        Bm25Shard* sh = nullptr;
        const int64_t nl = shard_row_count(G, s, n_docs);
        // a small index leaves the later shards empty (fewer documents than shards x RAG_SHARD_BLOCK): they exist,
        // hold nothing and answer with empty lists
        int rc = shard_build(&sh, s, nl, n_terms, tptr[s][(size_t)n_terms], tptr[s].data(), rows[s].data(), tfs[s].data(),
                             dls[s].data(), idf, avgdl, k1, b);
        if (rc != RAG_OK) { index_free(ix); return rc; }
        ix->sh.push_back(sh);
    }
    *out = ix;
    return RAG_OK;
}

int rag_bm25_destroy(rag_bm25_t* ix) {
    if (!ix) return RAG_OK;
    { IndexLock lk(ix); }
    index_free(ix);
    return RAG_OK;
}

int rag_bm25_info(const rag_bm25_t* ix, int* bytes_per_posting, int64_t* index_bytes) {
    if (!ix) return fail(RAG_EINVAL, "NULL argument");
    bool fast = g_bm25_fast != 0;
    int64_t bytes = 0;
    for (auto* s : ix->sh) {
        if (s->d.n_docs > 0) fast = fast && s->d.fast_ok;
        bytes += s->index_bytes;
    }
    if (bytes_per_posting) *bytes_per_posting = fast ? 4 : 12;
    if (index_bytes) *index_bytes = bytes;
    return RAG_OK;
}

int rag_bm25_query_bytes(const rag_bm25_t* ix, const int32_t* q_terms, const int32_t* q_ptr, int Q, int64_t* out_bytes,
                         int64_t* out_postings) {
    if (!ix || !q_ptr || Q <= 0 || (q_ptr[Q] > 0 && !q_terms)) return fail(RAG_EINVAL, "bad arguments");
    IndexLock lk(ix);
    for (int q = 0; q < Q; ++q) {
        int64_t bytes = 0, postings = 0;
        for (int i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {
            const int32_t t = q_terms[i];
            if (t < 0 || t >= ix->n_terms) continue;
            for (auto* s : ix->sh) {
                const int64_t df = s->h_term_ptr[(size_t)t + 1] - s->h_term_ptr[(size_t)t];
                if (df == 0 || s->h_idf[(size_t)t] == 0.0) continue;
                postings += df;
                const bool fast = g_bm25_fast && s->d.fast_ok && !s->h_cls.empty();
                if (!fast) bytes += df * 12;                                  // exact path: row id + fp64 impact
                else if (s->h_cls[(size_t)t] == kBmDense) bytes += s->d.n_docs * 2;   // one 16-bit column entry per row
                else bytes += df * 4;                                          // packed stream
            }
        }
        if (out_bytes) out_bytes[q] = bytes;
        if (out_postings) out_postings[q] = postings;
    }
    return RAG_OK;
}

int rag_bm25_search(rag_bm25_t* ix, const int32_t* q_terms, const int32_t* q_ptr, int Q, int k,
                    const uint8_t* allow_bitmap, int32_t* out_rows, double* out_scores, int32_t* out_counts) {
    RAG_TRY(require_init());
    if (!ix || !q_ptr || !out_rows || !out_scores || !out_counts || Q <= 0) return fail(RAG_EINVAL, "NULL argument");
    if (k <= 0 || k > RAG_MAX_K) return fail(RAG_ERANGE, "k=%d outside 1..%d", k, RAG_MAX_K);
    if (q_ptr[0] != 0 || q_ptr[Q] < 0 || (q_ptr[Q] > 0 && !q_terms)) return fail(RAG_EINVAL, "bad q_ptr / q_terms");
    IndexLock lk(ix);
    const int G = ix->n_shards;
    int max_tokens = 0;
    for (int q = 0; q < Q; ++q) max_tokens = std::max(max_tokens, q_ptr[q + 1] - q_ptr[q]);
    // a selective row filter (doc_filter keeping a few hundred chunks) is served from the listed rows alone
    std::vector<std::vector<int32_t>> listed(G);
    bool use_listed = false;
    if (allow_bitmap && g_bm25_rows_max > 0 && max_tokens <= 4096) {
        int64_t pop = 0;
        const int64_t nbytes = (ix->n_docs + 7) / 8;
        for (int64_t i = 0; i < nbytes && pop <= g_bm25_rows_max; ++i) pop += __builtin_popcount(allow_bitmap[i]);
        if (pop <= g_bm25_rows_max) {
            use_listed = true;
            for (int64_t i = 0; i < nbytes; ++i) {
                uint8_t v = allow_bitmap[i];
                while (v) {
                    const int bit = __builtin_ctz(v);
                    v &= (uint8_t)(v - 1);
                    const int64_t r = i * 8 + bit;
                    if (r < ix->n_docs) listed[shard_of_row(G, r)].push_back((int32_t)local_of_row(G, r));
                }
            }
        }
    }
    std::vector<SearchPlan> plans(G);
    std::vector<uint8_t> local_bm;
    const size_t rb = (size_t)Q * k * 4, sb = (size_t)Q * k * 8, cb = (size_t)Q * 4;
    for (int s = 0; s < G; ++s) {
        Bm25Shard& sh = *ix->sh[s];
        const uint8_t* bm = nullptr;
        if (allow_bitmap && !use_listed) {
            if (G == 1) bm = allow_bitmap;
            else {
                split_bitmap_rows(G, ix->n_docs, allow_bitmap, s, sh.d.n_docs, local_bm);
                bm = local_bm.data();              // copied into the shard's pinned block before the call returns
            }
        }
        RAG_TRY(shard_search_launch(sh, q_terms, q_ptr, Q, k, max_tokens, bm, use_listed ? listed[s].data() : nullptr,
                                    (int)listed[s].size(), plans[s]));
    }
    if (G == 1) {
        Bm25Shard& sh = *ix->sh[0];
        RAG_TRY(sh.cx.use());
        cudaStream_t st = sh.cx.stream();
        uint8_t* pin = reinterpret_cast<uint8_t*>(sh.cx.pinned);
        for (int attempt = 0; attempt < 2; ++attempt) {
            CU_TRY(cudaMemcpyAsync(pin, sh.outbuf.p, sb + rb + cb, cudaMemcpyDeviceToHost, st));   // scores | rows | counts
            CU_TRY(cudaStreamSynchronize(st));
            int n_redone = 0;
            if (attempt == 0)
                RAG_TRY(shard_search_redo(sh, Q, k, plans[0], reinterpret_cast<int32_t*>(pin + sb + rb), &n_redone));
            if (n_redone == 0) break;
        }
        memcpy(out_scores, pin, sb);
        memcpy(out_rows, pin + sb, rb);
        memcpy(out_counts, pin + sb + rb, cb);
        return RAG_OK;
    }
    // ---- sharded: redo flagged queries per shard, publish (score, GLOBAL id) into the primary shard's gather
    // buffers over peer memory, merge there, one D2H
    Bm25Shard& lead = *ix->sh[0];
    RAG_TRY(lead.cx.use());
    RAG_TRY(ix->g_scores.ensure((size_t)G * Q * k * 8));
    RAG_TRY(ix->g_ids.ensure((size_t)G * Q * k * 8));
    RAG_TRY(ix->m_scores.ensure(sb));
    RAG_TRY(ix->m_ids.ensure(sb));
    RAG_TRY(ix->m_counts.ensure(cb));
    for (int s = 0; s < G; ++s) {
        Bm25Shard& sh = *ix->sh[s];
        RAG_TRY(sh.cx.use());
        cudaStream_t st = sh.cx.stream();
        if (plans[s].fast) {
            int32_t* hc = reinterpret_cast<int32_t*>(sh.cx.pinned);
            CU_TRY(cudaMemcpyAsync(hc, sh.d_counts, cb, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            int n_redone = 0;
            RAG_TRY(shard_search_redo(sh, Q, k, plans[s], hc, &n_redone));
        }
        const int64_t n = (int64_t)Q * k;
        int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 4);
        bm25_publish_kernel<<<grid < 1 ? 1 : grid, 256, 0, st>>>(sh.d_rows, sh.d_scores, n, s, G,
                                                                   ix->g_ids.as<int64_t>() + (size_t)s * n,
                                                                   ix->g_scores.as<double>() + (size_t)s * n);
        CU_TRY(cudaGetLastError());
        ++R.n_launch;
        CU_TRY(cudaEventRecord(sh.cx.done, st));
    }
    RAG_TRY(lead.cx.use());
    cudaStream_t lst = lead.cx.stream();
    for (int s = 1; s < G; ++s) CU_TRY(cudaStreamWaitEvent(lst, ix->sh[s]->cx.done, 0));
    CU_TRY(merge_exact_launch(ix->g_scores.as<double>(), ix->g_ids.as<int64_t>(), G, Q, k, (int64_t)Q * k,
                              ix->m_scores.as<double>(), ix->m_ids.as<int64_t>(), ix->m_counts.as<int32_t>(), lst));
    ++R.n_launch;
    RAG_TRY(lead.cx.ensure_pinned(2 * sb + cb));
    uint8_t* pin = reinterpret_cast<uint8_t*>(lead.cx.pinned);
    CU_TRY(cudaMemcpyAsync(pin, ix->m_scores.p, sb, cudaMemcpyDeviceToHost, lst));
    CU_TRY(cudaMemcpyAsync(pin + sb, ix->m_ids.p, sb, cudaMemcpyDeviceToHost, lst));
    CU_TRY(cudaMemcpyAsync(pin + 2 * sb, ix->m_counts.p, cb, cudaMemcpyDeviceToHost, lst));
    CU_TRY(cudaStreamSynchronize(lst));
    memcpy(out_scores, pin, sb);
    const int64_t* ids = reinterpret_cast<const int64_t*>(pin + sb);
    for (size_t i = 0; i < (size_t)Q * k; ++i) out_rows[i] = (int32_t)ids[i];
    memcpy(out_counts, pin + 2 * sb, cb);
    return RAG_OK;
}

int rag_bm25_scores(rag_bm25_t* ix, const int32_t* q_terms, int n_q_terms, double* out_scores) {
    RAG_TRY(require_init());
    if (!ix || !out_scores || (n_q_terms > 0 && !q_terms)) return fail(RAG_EINVAL, "NULL argument");
    IndexLock lk(ix);
    const int G = ix->n_shards;
    std::vector<double> local;
    for (int s = 0; s < G; ++s) {
        Bm25Shard& sh = *ix->sh[s];
        RAG_TRY(sh.cx.use());
        cudaStream_t st = sh.cx.stream();
        std::vector<int64_t> ranges;
        int64_t total = 0;
        RAG_TRY(accumulate_query(sh, q_terms, n_q_terms, ranges, &total));
        double* dst = out_scores;
        if (G > 1) {
            local.resize((size_t)sh.d.n_docs);
            dst = local.data();
        }
        CU_TRY(cudaMemcpyAsync(dst, sh.d.score, (size_t)sh.d.n_docs * 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (G > 1)
            for (int64_t l = 0; l < sh.d.n_docs; ++l) out_scores[shard_global_row(l, s, G, RAG_SHARD_BLOCK)] = local[(size_t)l];
        const int n_ranges = (int)ranges.size() / 2;
        if (n_ranges > 0) {
            RAG_TRY(sh.ranges.ensure(ranges.size() * 8));
            CU_TRY(cudaMemcpyAsync(sh.ranges.p, ranges.data(), ranges.size() * 8, cudaMemcpyHostToDevice, st));
            CU_TRY(bm25_reset_launch(sh.d, sh.ranges.as<int64_t>(), n_ranges, bm25_harvest_grid(total, sh.cx.info->sm_count), st));
            ++R.n_launch;
            CU_TRY(cudaStreamSynchronize(st));
        }
    }
    return RAG_OK;
}

}  // extern "C"
