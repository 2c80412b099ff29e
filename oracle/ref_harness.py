"""ORACLE — test infrastructure, not product code.

Imports the reference's OWN Python (src/rag/retriever.py, src/rag/bm25_index.py)
unmodified from /root/reference, with oracle/rank_bm25.py standing in for the
absent third-party wheel.  Works only in the build container (the GPU box has
no /root/reference): used by oracle/gen_golden.py to produce tests/golden/ and
by the `needs_reference` tests.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("RAG_DPO_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "rag", "retriever.py"))


_cache = {}


def load():
    """Returns a dict with the reference modules: retriever, bm25_index."""
    if _cache:
        return _cache
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True          # the reference tree is read-only
    from . import rank_bm25 as restated
    sys.modules.setdefault("rank_bm25", restated)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _cache["retriever"] = importlib.import_module("src.rag.retriever")
    _cache["bm25_index"] = importlib.import_module("src.rag.bm25_index")
    return _cache


class FixedEmbeddingProvider:
    """embed(texts) -> the vectors registered for those texts (python lists, like
    src/utils/embedding_provider.py:118-147 returns)."""

    def __init__(self, table):
        self.table = table

    def embed(self, texts):
        return [list(map(float, self.table[t])) for t in texts]


class FixedQueryExpander:
    """Stands in for src/rag/query_expander.py:66-113 (an LLM call): returns
    [query] + the registered reformulations."""

    def __init__(self, table):
        self.table = table

    def expand(self, query):
        return [query] + list(self.table.get(query, []))
