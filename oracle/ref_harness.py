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
    """Returns a dict with the reference modules: retriever, bm25_index, reranker."""
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
    _cache["reranker"] = importlib.import_module("src.rag.reranker")     # its model is loaded lazily: never here
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


class TableScorer:
    """Stands in for sentence_transformers.CrossEncoder (src/rag/reranker.py:74-107, model inference): predict(pairs)
    returns the registered fp32 score of every (query, text) pair, as a float32 array like the real model."""

    def __init__(self, table):
        self.table = table
        self.calls = []

    def predict(self, pairs, batch_size=32, show_progress_bar=False):
        import numpy as np
        self.calls.append([list(p) for p in pairs])
        return np.array([self.table[(q, t)] for q, t in pairs], dtype=np.float32)


class TagTopicMatcher:
    """Stands in for src/utils/rgpd_topics.py TopicMatcher.topic_boost (an embedding model behind it): 0.15 for an
    exact topic/tag match (the reference's own shortcut, rgpd_topics.py:206-209), else a registered value."""

    def __init__(self, partial=None):
        self.partial = partial or {}

    def topic_boost(self, question_topics, chunk_tags_str, threshold=0.65):
        if not question_topics or not chunk_tags_str:
            return 0.0
        tags = [t.strip() for t in chunk_tags_str.split(",") if t.strip()]
        best = 0.0
        for topic in question_topics:
            for tag in tags:
                if topic.lower() == tag.lower():
                    return 0.15
                best = max(best, self.partial.get((topic, tag), 0.0))
        return best
