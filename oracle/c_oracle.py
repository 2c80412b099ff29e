"""ORACLE — test infrastructure, not product code.
numpy-friendly wrappers over oracle.c (liboracle.so)."""
import numpy as np

from . import lib


def _p(a):
    return a.ctypes.data if a is not None else None


def dense_topk(q32, rows, dtype, k, allow_bitmap=None):
    """q32 (B,D) fp32; rows (n,D) in storage dtype (fp32 array, or uint16 bit
    patterns for bf16/fp16).  Returns rows int64 (B,k) [-1 padded], scores fp64
    (B,k), counts int32 (B,)."""
    q32 = np.ascontiguousarray(np.atleast_2d(q32), dtype=np.float32)
    rows = np.ascontiguousarray(rows)
    B, d = q32.shape
    n = rows.shape[0]
    out_r = np.empty((B, k), dtype=np.int64)
    out_s = np.empty((B, k), dtype=np.float64)
    out_c = np.empty(B, dtype=np.int32)
    ab = np.ascontiguousarray(allow_bitmap, dtype=np.uint8) if allow_bitmap is not None else None
    lib().orc_dense_topk(_p(q32), B, _p(rows), dtype, n, d, k, _p(ab), _p(out_r), _p(out_s), _p(out_c))
    return out_r, out_s, out_c


def dense_scores(q32, rows, dtype):
    q32 = np.ascontiguousarray(q32, dtype=np.float32)
    rows = np.ascontiguousarray(rows)
    n, d = rows.shape
    out = np.empty(n, dtype=np.float64)
    lib().orc_dense_scores(_p(q32), _p(rows), dtype, n, d, _p(out))
    return out


def bm25_scores(term_ptr, post_row, post_tf, doc_len, idf, avgdl, k1, b, q_terms):
    term_ptr = np.ascontiguousarray(term_ptr, dtype=np.int64)
    post_row = np.ascontiguousarray(post_row, dtype=np.int32)
    post_tf = np.ascontiguousarray(post_tf, dtype=np.int32)
    doc_len = np.ascontiguousarray(doc_len, dtype=np.int32)
    idf = np.ascontiguousarray(idf, dtype=np.float64)
    q_terms = np.ascontiguousarray(q_terms, dtype=np.int32)
    n = len(doc_len)
    out = np.empty(n, dtype=np.float64)
    lib().orc_bm25_scores(_p(term_ptr), _p(post_row), _p(post_tf), _p(doc_len), _p(idf), avgdl, k1, b,
                          len(term_ptr) - 1, _p(q_terms), len(q_terms), n, _p(out))
    return out


def bm25_select(score, k, allow_bitmap=None):
    score = np.ascontiguousarray(score, dtype=np.float64)
    out_r = np.empty(k, dtype=np.int64)
    out_s = np.empty(k, dtype=np.float64)
    ab = np.ascontiguousarray(allow_bitmap, dtype=np.uint8) if allow_bitmap is not None else None
    c = lib().orc_bm25_select(_p(score), len(score), _p(ab), k, _p(out_r), _p(out_s))
    return out_r[:c].copy(), out_s[:c].copy()


def rrf(ids, weights, rrf_k=60, top=None):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    R, L = ids.shape
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    top = R * L if top is None else top
    out_i = np.empty(max(top, 1), dtype=np.int32)
    out_s = np.empty(max(top, 1), dtype=np.float64)
    c = lib().orc_rrf(_p(ids), _p(weights), R, L, rrf_k, top, _p(out_i), _p(out_s))
    return out_i[:c].copy(), out_s[:c].copy()


class HnswIndex:
    """Restated hnswlib-style index (oracle/hnsw.c): what chromadb 1.4.1 builds for a cosine collection with its
    defaults M=16, ef_construction=100, ef_search=100.  PARITY UNPINNED (statistical twin); used only to report
    recall@k of the approximate store against the exact search."""

    def __init__(self, x_unit, M=16, ef_construction=100, seed=100):
        self.x = np.ascontiguousarray(x_unit, dtype=np.float32)
        n, d = self.x.shape
        assert d % 16 == 0
        self._h = lib().hnsw_build(_p(self.x), n, d, M, ef_construction, seed)

    def query(self, q_unit, k, ef_search=100):
        q = np.ascontiguousarray(np.atleast_2d(q_unit), dtype=np.float32)
        ids = np.full((len(q), k), -1, dtype=np.int32)
        dist = np.zeros((len(q), k), dtype=np.float32)
        for i in range(len(q)):
            lib().hnsw_search(self._h, _p(q[i]), k, ef_search, _p(ids[i]), _p(dist[i]))
        return ids, dist

    def close(self):
        if self._h:
            lib().hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
