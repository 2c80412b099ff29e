"""Reranker input/output step on the device — the part of CrossEncoderReranker.rerank
(src/rag/reranker.py:109-227) that is not model inference: it is called right after
retrieve_candidates (src/rag/pipeline.py:244-256) with ~40 candidates per question.

    pairs   (query, "heading\\ntext"[:max_length*4])                       src/rag/reranker.py:137-145
    scores  <- scorer.predict(pairs)     (the cross-encoder: OUTSIDE this package, injected)
    final   float(score) + topic_boost, stable sort descending, [:top_k],
            drop < min_score, never fewer than 3                           src/rag/reranker.py:172-211

`DeviceRerankStep.rerank` has the reference's signature and return type; `rerank_batch` serves many
questions with ONE scorer call and ONE rag_rerank_select launch (the order / cut / threshold run in
csrc/rrf.cu: rerank_select_kernel).  There is no CPU path for the select: without the library it raises.
"""
import logging
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

logger = logging.getLogger(__name__)

RERANK_MAX_CANDIDATES = 1024      # csrc/rrf.cu kRerankMax


@dataclass
class RankedChunk:
    """same fields as src/rag/reranker.py:26-34"""
    chunk_id: str
    text: str
    document_path: str
    rerank_score: float
    original_rank: int
    metadata: dict


def rerank_select(scores, boosts=None, lens=None, top_k=8, min_score=0.08):
    """scores (Q, L) fp32 model scores in candidate order; boosts (Q, L) fp64 or None; lens (Q,) or None.
    Returns idx (Q, top_k) int32 positions in the candidate list [-1 padded], final scores (Q, top_k) fp64,
    counts (Q,)."""
    scores = np.ascontiguousarray(np.atleast_2d(scores), dtype=np.float32)
    Q, L = scores.shape
    if top_k < 3:
        raise ValueError("top_k below 3: use DeviceRerankStep, which applies the reference's keep-3 rule on the host")
    b = None if boosts is None else np.ascontiguousarray(np.atleast_2d(boosts), dtype=np.float64)
    ln = None if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
    idx = np.empty((Q, top_k), dtype=np.int32)
    out = np.empty((Q, top_k), dtype=np.float64)
    counts = np.empty(Q, dtype=np.int32)
    _lib.check(_lib.lib().rag_rerank_select(_lib.ptr(scores), _lib.ptr(b), _lib.ptr(ln), Q, L, int(top_k), float(min_score),
                                            _lib.ptr(idx), _lib.ptr(out), _lib.ptr(counts)))
    return idx, out, counts


class DeviceRerankStep:
    """Drop-in for CrossEncoderReranker (src/rag/reranker.py:37-231) around an injected scorer.

    scorer: object with predict(pairs, batch_size=..., show_progress_bar=False) -> sequence of floats
    (sentence_transformers.CrossEncoder has exactly this method; the reference builds one lazily)."""

    def __init__(self, scorer, batch_size: int = 32, max_length: int = 512, min_score: float = 0.08):
        self._model = scorer
        self.batch_size = batch_size
        self.max_length = max_length
        self.min_score = min_score
        self._is_loaded = True

    @property
    def is_loaded(self) -> bool:
        return self._is_loaded

    # -- the input side: src/rag/reranker.py:137-145
    def _pairs(self, query, chunks):
        pairs = []
        for chunk in chunks:
            text = chunk.text
            heading = chunk.metadata.get("heading", "")
            if heading:
                text = f"{heading}\n{text}"
            pairs.append((query, text[:self.max_length * 4]))
        return pairs

    @staticmethod
    def _fallback(chunks, top_k):
        # scorer failure: the original order, similarity as the score (src/rag/reranker.py:154-166)
        return [RankedChunk(chunk_id=c.chunk_id, text=c.text, document_path=c.document_path,
                            rerank_score=c.similarity_score, original_rank=i, metadata=c.metadata)
                for i, c in enumerate(chunks[:top_k])]

    def rerank(self, query: str, chunks: List, top_k: int = 8, topic_matcher=None,
               question_topics: Optional[List[str]] = None) -> List[RankedChunk]:
        return self.rerank_batch([query], [chunks], top_k, topic_matcher, [question_topics])[0]

    def rerank_batch(self, queries: Sequence[str], chunk_lists: Sequence[List], top_k: int = 8, topic_matcher=None,
                     question_topics_list: Optional[Sequence[Optional[List[str]]]] = None) -> List[List[RankedChunk]]:
        n_q = len(queries)
        topics = list(question_topics_list) if question_topics_list is not None else [None] * n_q
        results: List[Optional[List[RankedChunk]]] = [None] * n_q
        live = [i for i in range(n_q) if chunk_lists[i]]
        for i in range(n_q):
            if not chunk_lists[i]:
                results[i] = []
        if not live:
            return results
        pairs, spans = [], []
        for i in live:
            p = self._pairs(queries[i], chunk_lists[i])
            spans.append((len(pairs), len(p)))
            pairs.extend(p)
        try:
            flat = self._model.predict(pairs, batch_size=self.batch_size, show_progress_bar=False)
        except Exception as e:
            logger.error(f"rerank scorer failed: {e}")
            for i in live:
                results[i] = self._fallback(chunk_lists[i], top_k)
            return results
        L = max(n for _, n in spans)
        if L > RERANK_MAX_CANDIDATES:
            raise ValueError(f"{L} candidates for one question exceed {RERANK_MAX_CANDIDATES}")
        scores = np.zeros((len(live), L), dtype=np.float32)
        boosts = np.zeros((len(live), L), dtype=np.float64)
        lens = np.zeros(len(live), dtype=np.int32)
        any_boost = False
        for row, (i, (off, n)) in enumerate(zip(live, spans)):
            scores[row, :n] = np.asarray(flat[off:off + n], dtype=np.float32)
            lens[row] = n
            if topic_matcher is not None and topics[i]:
                for j, chunk in enumerate(chunk_lists[i]):
                    boost = topic_matcher.topic_boost(topics[i], chunk.metadata.get("rgpd_topics", ""))
                    if boost > 0:
                        boosts[row, j] = boost
                        any_boost = True
        k_dev = max(int(top_k), 3)
        # top_k < 3: the reference still returns ranked[:3] when fewer than 3 pass — take the first 3 unfiltered and
        # apply its rule below
        min_dev = self.min_score if top_k >= 3 else -np.inf
        idx, final, counts = rerank_select(scores, boosts if any_boost else None, lens, k_dev, min_dev)
        for row, i in enumerate(live):
            chunks = chunk_lists[i]
            ranked = [RankedChunk(chunk_id=chunks[j].chunk_id, text=chunks[j].text, document_path=chunks[j].document_path,
                                  rerank_score=float(final[row, r]), original_rank=int(j), metadata=chunks[j].metadata)
                      for r, j in enumerate(idx[row, :counts[row]])]
            if top_k < 3:
                kept = [r for r in ranked[:top_k] if r.rerank_score >= self.min_score]
                ranked = ranked[:3] if (len(kept) < 3 and len(chunks) >= 3) else kept
            if not ranked:
                # fewer than 3 candidates, all below min_score: the reference's closing log line indexes the empty
                # result and the call raises (src/rag/reranker.py:221-225); same error here
                raise IndexError("list index out of range")
            results[i] = ranked
        return results
