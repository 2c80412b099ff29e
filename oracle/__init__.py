"""ORACLE — test infrastructure, not product code.

CPU restatement of the retrieval hot path of MatJoss/RAG-DPO used ONLY as the
checker: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this package; nothing under
rag-dpo_b200/ does.

Parity status:
  * reference-owned code (RRF, tokenizer, BM25 select, fusion orchestration) is
    PINNED: tests/golden/*.json were produced by running the reference's own
    unmodified Python (oracle/gen_golden.py via oracle/ref_harness.py).
  * the third-party arithmetic the reference delegates to — chromadb==1.4.1
    (dense cosine kNN) and rank-bm25==0.2.2 (BM25Okapi) — is absent from
    /root/reference, not installed and not installable: PARITY UNPINNED for those
    two; they are restated from their published algorithms.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle.c + hnsw.c -> liboracle.so with gcc (no reference sources involved)."""
    so = os.path.join(_HERE, "liboracle.so")
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.c", "hnsw.c", "Makefile"))
    if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    """ctypes handle on the C restatement."""
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c = ctypes
        L.orc_dense_score.restype = c.c_double
        L.orc_dense_score.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int64, c.c_int]
        L.orc_dense_scores.restype = None
        L.orc_dense_scores.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int64, c.c_int, c.c_void_p]
        L.orc_dense_topk.restype = None
        L.orc_dense_topk.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_int64, c.c_int, c.c_int,
                                     c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
        L.orc_bm25_scores.restype = None
        L.orc_bm25_scores.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                      c.c_double, c.c_double, c.c_double, c.c_int64, c.c_void_p, c.c_int,
                                      c.c_int64, c.c_void_p]
        L.orc_bm25_select.restype = c.c_int
        L.orc_bm25_select.argtypes = [c.c_void_p, c.c_int64, c.c_void_p, c.c_int, c.c_void_p, c.c_void_p]
        L.orc_rrf.restype = c.c_int
        L.orc_rrf.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p]
        L.hnsw_build.restype = c.c_void_p
        L.hnsw_build.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_uint64]
        L.hnsw_search.restype = c.c_int
        L.hnsw_search.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_void_p]
        L.hnsw_free.restype = None
        L.hnsw_free.argtypes = [c.c_void_p]
        _LIB = L
    return _LIB
