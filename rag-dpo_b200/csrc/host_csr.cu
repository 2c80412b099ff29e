// host_csr.cu — host-side (multi-threaded C++) construction of the CSR postings that rag_bm25_create uploads.
//
// Replaces the per-document dict building of rank_bm25.BM25Okapi._initialize (rank-bm25 0.2.2, called from
// ChunkBM25Index.build_from_collection, src/rag/bm25_index.py:236, and SummaryBM25Index.build, :126): for every
// document the term frequencies, for every term the ascending list of (row, tf).  No GPU is needed for this call.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/b200rag.h"

namespace {

struct Part {
    int64_t d0 = 0, d1 = 0;            // documents of this worker
    std::vector<int32_t> term, tf;     // (term, tf) pairs in document order, terms ascending inside a document
    std::vector<int32_t> ucount;       // distinct terms per document
    std::vector<int64_t> df;           // per-term document frequency inside this part
};

}  // namespace

extern "C" int rag_csr_build(int64_t n_docs, const int64_t* doc_ptr, const int32_t* tokens, int64_t n_terms,
                             int64_t* term_ptr, int32_t* post_row, int32_t* post_tf, int64_t capacity, int64_t* nnz_out,
                             int n_threads) {
    if (n_docs < 0 || n_terms < 0 || !doc_ptr || !term_ptr || !nnz_out || n_docs > 0x7FFFFFF0LL) return RAG_EINVAL;
    const int64_t total = doc_ptr[n_docs];
    if (total > 0 && !tokens) return RAG_EINVAL;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    if ((int64_t)T > n_docs) T = (int)std::max<int64_t>(1, n_docs);
    std::vector<Part> parts(T);
    // split the documents into T contiguous parts of about the same number of tokens
    {
        int64_t d = 0;
        for (int t = 0; t < T; ++t) {
            parts[t].d0 = d;
            const int64_t want = total * (t + 1) / T;
            while (d < n_docs && (doc_ptr[d + 1] <= want || t == T - 1)) ++d;
            if (t == T - 1) d = n_docs;
            parts[t].d1 = d;
        }
    }
    std::atomic<bool> bad{false};
    auto phase1 = [&](int t) {
        Part& p = parts[t];
        p.df.assign((size_t)n_terms, 0);
        p.ucount.reserve((size_t)(p.d1 - p.d0));
        const int64_t ntok = doc_ptr[p.d1] - doc_ptr[p.d0];
        p.term.reserve((size_t)ntok);
        p.tf.reserve((size_t)ntok);
        std::vector<int32_t> buf;
        for (int64_t d = p.d0; d < p.d1; ++d) {
            const int64_t a = doc_ptr[d], b = doc_ptr[d + 1];
            buf.assign(tokens + a, tokens + b);
            std::sort(buf.begin(), buf.end());
            int32_t u = 0;
            for (size_t i = 0; i < buf.size();) {
                size_t j = i;
                while (j < buf.size() && buf[j] == buf[i]) ++j;
                const int32_t w = buf[i];
                if (w < 0 || w >= n_terms) { bad = true; return; }
                p.term.push_back(w);
                p.tf.push_back((int32_t)(j - i));
                ++p.df[(size_t)w];
                ++u;
                i = j;
            }
            p.ucount.push_back(u);
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(phase1, t);
        phase1(0);
        for (auto& x : th) x.join();
    }
    if (bad.load()) return RAG_EINVAL;
    // term_ptr = prefix sum of the document frequencies; per-part write cursors (parts are in row order, so the
    // rows of a term ascend)
    term_ptr[0] = 0;
    for (int64_t w = 0; w < n_terms; ++w) {
        int64_t s = 0;
        for (int t = 0; t < T; ++t) s += parts[t].df[(size_t)w];
        term_ptr[w + 1] = term_ptr[w] + s;
    }
    const int64_t nnz = term_ptr[n_terms];
    *nnz_out = nnz;
    if (nnz > capacity || (nnz > 0 && (!post_row || !post_tf))) return RAG_ERANGE;
    for (int64_t w = 0; w < n_terms; ++w) {
        int64_t cur = term_ptr[w];
        for (int t = 0; t < T; ++t) {
            const int64_t c = parts[t].df[(size_t)w];
            parts[t].df[(size_t)w] = cur;              // becomes the part's write cursor for this term
            cur += c;
        }
    }
    auto phase2 = [&](int t) {
        Part& p = parts[t];
        size_t i = 0;
        for (int64_t d = p.d0; d < p.d1; ++d) {
            const int32_t u = p.ucount[(size_t)(d - p.d0)];
            for (int32_t j = 0; j < u; ++j, ++i) {
                const int64_t pos = p.df[(size_t)p.term[i]]++;
                post_row[pos] = (int32_t)d;
                post_tf[pos] = p.tf[i];
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(phase2, t);
        phase2(0);
        for (auto& x : th) x.join();
    }
    return RAG_OK;
}
