/* ORACLE — test infrastructure, not product code.
 *
 * Plain-C restatement of the arithmetic on RAG-DPO's retrieval hot path
 * (citations relative to /root/reference):
 *
 *   orc_dense_*   collection.query(...) as issued at src/rag/retriever.py:215-220
 *                 and :380-385 on a cosine-space collection
 *                 (src/processing/create_chromadb_index.py:100-106): score =
 *                 <q^,x^>, distance = 1 - score, ascending distance, ties ->
 *                 lowest row, rows failing the filter excluded before selection.
 *                 chromadb==1.4.1 (requirements.txt:33) is absent: PARITY UNPINNED
 *                 for the third-party HNSW itself; this is the exact search it
 *                 approximates.
 *   orc_bm25_*    rank-bm25==0.2.2 BM25Okapi.get_scores (requirements.txt:38,
 *                 called at src/rag/bm25_index.py:153,265) + the reference-owned
 *                 select at src/rag/bm25_index.py:267-279.  PARITY UNPINNED for
 *                 the third-party arithmetic (see oracle/rank_bm25.py).
 *   orc_rrf       reciprocal_rank_fusion, src/rag/retriever.py:66-90, followed
 *                 by the stable descending sort of :464-465.
 *
 * Compile with -ffp-contract=off: every operation must round once, like numpy.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

static inline float bf16_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t u;
    if (exp == 0) {
        if (man == 0) {
            u = sign;
        } else {                       /* subnormal: renormalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3FFu;
            u = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        u = sign | 0x7F800000u | (man << 13);
    } else {
        u = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline double load_elem(const void* rows, int dtype, int64_t idx) {
    switch (dtype) {
        case DT_F32: return (double)((const float*)rows)[idx];
        case DT_BF16: return (double)bf16_to_f32(((const uint16_t*)rows)[idx]);
        default: return (double)f16_to_f32(((const uint16_t*)rows)[idx]);
    }
}

/* Canonical score: 32 fp64 partial sums — p[l] owns the 8-element
 * groups l, l+32, l+64, ... and adds their products in ascending element order —
 * then a fixed halving tree.  The products are exact in fp64 (24-bit x <=24-bit). */
double orc_dense_score(const float* q, const void* rows, int dtype, int64_t row, int d) {
    double p[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.0;
    for (int g = 0; g < d / 8; ++g)
        for (int i = 0; i < 8; ++i)
            p[g & 31] += (double)q[8 * g + i] * load_elem(rows, dtype, row * (int64_t)d + 8 * g + i);
    for (int off = 16; off >= 1; off >>= 1)
        for (int l = 0; l < off; ++l) p[l] += p[l + off];
    return p[0];
}

void orc_dense_scores(const float* q, const void* rows, int dtype, int64_t n, int d, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) out[r] = orc_dense_score(q, rows, dtype, r, d);
}

/* better(a,b): a ranks before b under (score desc, row asc) */
static inline int better(double sa, int64_t ra, double sb, int64_t rb) {
    return sa > sb || (sa == sb && ra < rb);
}

/* Exact top-k for nq queries.  allow: optional bitmap (bit r of byte r/8), NULL = all.
 * out_rows/out_scores: nq*k, padded with -1 / 0; out_counts: nq. */
void orc_dense_topk(const float* q, int nq, const void* rows, int dtype, int64_t n, int d, int k,
                    const uint8_t* allow, int64_t* out_rows, double* out_scores, int32_t* out_counts) {
    double* sc = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int b = 0; b < nq; ++b) {
        orc_dense_scores(q + (int64_t)b * d, rows, dtype, n, d, sc);
        int cnt = 0;
        int64_t* orow = out_rows + (int64_t)b * k;
        double* osc = out_scores + (int64_t)b * k;
        for (int64_t r = 0; r < n; ++r) {
            if (allow && !((allow[r >> 3] >> (r & 7)) & 1)) continue;
            if (cnt == k && !better(sc[r], r, osc[k - 1], orow[k - 1])) continue;
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && better(sc[r], r, osc[pos - 1], orow[pos - 1])) {
                osc[pos] = osc[pos - 1];
                orow[pos] = orow[pos - 1];
                --pos;
            }
            osc[pos] = sc[r];
            orow[pos] = r;
            if (cnt < k) ++cnt;
        }
        for (int i = cnt; i < k; ++i) { orow[i] = -1; osc[i] = 0.0; }
        out_counts[b] = cnt;
    }
    free(sc);
}

/* BM25Okapi.get_scores over CSR postings.  q_terms may repeat (each repeat adds
 * again) and may hold -1 for tokens outside the vocabulary (adds nothing). */
void orc_bm25_scores(const int64_t* term_ptr, const int32_t* post_row, const int32_t* post_tf,
                     const int32_t* doc_len, const double* idf, double avgdl, double k1, double b,
                     int64_t n_terms, const int32_t* q_terms, int nq_terms, int64_t n, double* score) {
    for (int64_t r = 0; r < n; ++r) score[r] = 0.0;
    for (int i = 0; i < nq_terms; ++i) {
        int32_t t = q_terms[i];
        if (t < 0 || t >= n_terms) continue;
        double w = idf[t];
        if (w == 0.0) continue;
        for (int64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
            int32_t r = post_row[p];
            double tf = (double)post_tf[p];
            double num = tf * (k1 + 1.0);
            double len_norm = b * (double)doc_len[r];
            len_norm = len_norm / avgdl;
            double inner = (1.0 - b) + len_norm;
            double den = tf + k1 * inner;
            double frac = num / den;
            score[r] += w * frac;
        }
    }
}

/* select of src/rag/bm25_index.py:267-279: score > 0, filter, stable sort desc, top_k */
int orc_bm25_select(const double* score, int64_t n, const uint8_t* allow, int k,
                    int64_t* out_rows, double* out_scores) {
    int cnt = 0;
    for (int64_t r = 0; r < n; ++r) {
        if (!(score[r] > 0.0)) continue;
        if (allow && !((allow[r >> 3] >> (r & 7)) & 1)) continue;
        if (cnt == k && !better(score[r], r, out_scores[k - 1], out_rows[k - 1])) continue;
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && better(score[r], r, out_scores[pos - 1], out_rows[pos - 1])) {
            out_scores[pos] = out_scores[pos - 1];
            out_rows[pos] = out_rows[pos - 1];
            --pos;
        }
        out_scores[pos] = score[r];
        out_rows[pos] = r;
        if (cnt < k) ++cnt;
    }
    return cnt;
}

/* Weighted RRF over integer ids.  ids: R x L (row-major), negative = padding.
 * Output: distinct ids ordered by (fused score desc, first-seen position asc),
 * at most `top`.  Returns the number written. */
int orc_rrf(const int32_t* ids, const double* weights, int R, int L, int rrf_k, int top,
            int32_t* out_ids, double* out_scores) {
    int cap = R * L;
    int32_t* uid = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cap > 0 ? cap : 1));
    double* usc = (double*)malloc(sizeof(double) * (size_t)(cap > 0 ? cap : 1));
    int nu = 0;
    for (int r = 0; r < R; ++r) {
        int rank = 0;                      /* enumerate() index among non-padding entries */
        for (int j = 0; j < L; ++j) {
            int32_t id = ids[r * L + j];
            if (id < 0) continue;
            double add = weights[r] / (double)(rrf_k + rank + 1);
            ++rank;
            int u = 0;
            while (u < nu && uid[u] != id) ++u;
            if (u == nu) { uid[nu] = id; usc[nu] = 0.0; ++nu; }
            usc[u] += add;
        }
    }
    /* stable insertion sort, descending */
    int* ord = (int*)malloc(sizeof(int) * (size_t)(nu > 0 ? nu : 1));
    for (int i = 0; i < nu; ++i) {
        int pos = i;
        while (pos > 0 && usc[ord[pos - 1]] < usc[i]) { ord[pos] = ord[pos - 1]; --pos; }
        ord[pos] = i;
    }
    int nout = nu < top ? nu : top;
    for (int i = 0; i < nout; ++i) { out_ids[i] = uid[ord[i]]; out_scores[i] = usc[ord[i]]; }
    free(uid); free(usc); free(ord);
    return nout;
}
