"""Weighted Reciprocal Rank Fusion on the device — same contract as
reciprocal_rank_fusion(rankings, k=60, weights=None) -> Dict[id, score]
(src/rag/retriever.py:66-90), plus the fused order of the fusion tail
(src/rag/retriever.py:464-467: stable descending sort, first-seen order).
"""
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


RRF_MAX_ENTRIES = 8192      # R * L of one question (csrc/rrf.cu kRrfMaxEntries)


def rrf_capacity():
    return RRF_MAX_ENTRIES


def rrf_fuse_rows(ids, weights, rrf_k=60, top=None):
    """ids: int32 array (Q, R, L), negative = padding; weights (Q, R) or (R,).
    Returns out_ids (Q, top) [-1 padded], out_scores (Q, top), counts (Q,)."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    if ids.ndim == 2:
        ids = ids[None]
    Q, R, L = ids.shape
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim == 1:
        w = np.broadcast_to(w, (Q, R))
    w = np.ascontiguousarray(w, dtype=np.float64)
    top = R * L if top is None else int(top)
    out_ids = np.empty((Q, top), dtype=np.int32)
    out_scores = np.empty((Q, top), dtype=np.float64)
    counts = np.empty(Q, dtype=np.int32)
    _lib.check(_lib.lib().rag_rrf_fuse(_lib.ptr(ids), _lib.ptr(w), Q, R, L, int(rrf_k), top, _lib.ptr(out_ids),
                                       _lib.ptr(out_scores), _lib.ptr(counts)))
    return out_ids, out_scores, counts


def fuse_ranked(rankings: Sequence[Sequence[Hashable]], k: int = 60,
                weights: Optional[List[float]] = None, top: Optional[int] = None) -> Tuple[list, List[float]]:
    """Fused ids in final order + their scores."""
    if weights is None:
        weights = [1.0] * len(rankings)
    R = min(len(rankings), len(weights))           # zip() semantics of the reference
    if R == 0:
        return [], []
    L = max(1, max(len(r) for r in rankings[:R]))
    code: Dict[Hashable, int] = {}
    names = []
    arr = np.full((1, R, L), -1, dtype=np.int32)
    for r in range(R):
        for j, key in enumerate(rankings[r]):
            c = code.get(key)
            if c is None:
                c = code[key] = len(names)
                names.append(key)
            arr[0, r, j] = c
    out_ids, out_scores, counts = rrf_fuse_rows(arr, np.asarray(weights[:R], dtype=np.float64), k, top)
    n = int(counts[0])
    return [names[i] for i in out_ids[0, :n]], out_scores[0, :n].tolist()


def reciprocal_rank_fusion(rankings: List[List[str]], k: int = 60,
                           weights: Optional[List[float]] = None) -> Dict[str, float]:
    """Same signature and return value as the reference function."""
    ids, scores = fuse_ranked(rankings, k=k, weights=weights)
    # the reference's dict is in first-seen order
    order = {}
    for ranking in rankings[:len(weights) if weights is not None else len(rankings)]:
        for key in ranking:
            order.setdefault(key, len(order))
    fused = dict(zip(ids, scores))
    return {key: fused[key] for key in order}
