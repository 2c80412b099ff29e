"""HybridRetriever — host-side mirror of RAGRetriever
(src/rag/retriever.py:93-578): same constructor arguments, same
retrieve() / retrieve_candidates() semantics and return types, so that the
parity tests read like calls into the reference.  It exists because the
reference package itself cannot travel to the GPU box; in a real deployment
the reference's own RAGRetriever is constructed with DeviceCollection and
DeviceChunkBM25Index instead.

All arithmetic on the path (similarity, top-k, BM25, RRF) runs on the device
through the injected objects; this file only orchestrates.
"""
import logging
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

from . import rrf as _rrf

logger = logging.getLogger(__name__)


@dataclass
class RetrievedChunk:
    """mirror of src/rag/retriever.py:22-42"""
    chunk_id: str
    text: str
    document_path: str
    chunk_nature: str
    chunk_index: int
    confidence: str
    distance: float
    metadata: Dict[str, Any]
    bm25_score: float = 0.0
    semantic_score: float = 0.0
    hybrid_score: float = 0.0

    @property
    def similarity_score(self) -> float:
        return 1.0 / (1.0 + self.distance)


@dataclass
class RetrievedDocument:
    """mirror of src/rag/retriever.py:45-63"""
    document_path: str
    chunks: List[RetrievedChunk]
    avg_similarity: float
    primary_nature: str

    def __post_init__(self):
        if self.chunks:
            self.avg_similarity = sum(c.similarity_score for c in self.chunks) / len(self.chunks)
            natures = [c.chunk_nature for c in self.chunks]
            self.primary_nature = max(set(natures), key=natures.count)
        else:
            self.avg_similarity = 0.0
            self.primary_nature = "UNKNOWN"


def _chunk_from_meta(chunk_id, text, meta, distance):
    return RetrievedChunk(chunk_id=chunk_id, text=text, document_path=meta.get("document_path", ""),
                          chunk_nature=meta.get("chunk_nature", "UNKNOWN"), chunk_index=meta.get("chunk_index", 0),
                          confidence=meta.get("confidence", "unknown"), distance=distance, metadata=meta)


class HybridRetriever:
    def __init__(self, collection, llm_provider=None, embedding_provider=None, summary_bm25_index=None,
                 chunk_bm25_index=None, query_expander=None, n_documents: int = 5, n_chunks_per_doc: int = 3,
                 fetch_multiplier: int = 10, summary_prefilter_k: int = 20, enable_hybrid: bool = True,
                 enable_summary_prefilter: bool = True,
                 acronym_expander: Optional[Callable[[str], str]] = None,
                 fuse: Callable = _rrf.reciprocal_rank_fusion):
        self.collection = collection
        self.llm_provider = llm_provider
        self.embedding_provider = embedding_provider
        self.summary_bm25 = summary_bm25_index
        self.chunk_bm25 = chunk_bm25_index
        self.query_expander = query_expander
        self.n_documents = n_documents
        self.n_chunks_per_doc = n_chunks_per_doc
        self.fetch_multiplier = fetch_multiplier
        self.summary_prefilter_k = summary_prefilter_k
        self.enable_hybrid = enable_hybrid
        self.enable_summary_prefilter = enable_summary_prefilter
        # src/utils/acronyms.py:151-198 is host string prep upstream of the boundary; inject it
        self.acronym_expander = acronym_expander or (lambda s: s)
        self.fuse = fuse

    def _embed(self, texts):
        if self.embedding_provider is not None:
            return self.embedding_provider.embed(texts)
        return self.llm_provider.embed(texts)

    # ---- shared loop (retriever.py:207-290 and :372-452) --------------------
    def _gather(self, query, where_filter, n_fetch, backfill, bm25_all_queries):
        expanded = self.acronym_expander(query)
        all_queries = self.query_expander.expand(expanded) if self.query_expander is not None else [expanded]
        doc_filter = None
        if self.enable_summary_prefilter and self.summary_bm25 is not None and self.summary_bm25._is_built:
            doc_filter = self.summary_bm25.get_relevant_doc_paths(expanded, top_k=self.summary_prefilter_k)
        rankings, weights = [], []
        chunk_map: Dict[str, RetrievedChunk] = {}
        for q_idx, q in enumerate(all_queries):
            q_weight = 2.0 if q_idx == 0 else 1.0
            emb = self._embed([q])[0]
            try:
                res = self.collection.query(query_embeddings=[emb], n_results=n_fetch, where=where_filter,
                                            include=["documents", "metadatas", "distances"])
            except Exception as e:      # the reference logs and skips this query variant (retriever.py:221-223)
                logger.error("dense query failed (variant %d): %s", q_idx, e)
                continue
            chunks = [_chunk_from_meta(i, t, m, d) for i, t, m, d in
                      zip(res["ids"][0], res["documents"][0], res["metadatas"][0], res["distances"][0])]
            if doc_filter:
                kept = [c for c in chunks if c.document_path in doc_filter]
                if len(kept) < backfill:
                    kept.extend([c for c in chunks if c not in kept][:backfill - len(kept)])
                chunks = kept
            for c in chunks:
                c.semantic_score = c.similarity_score
            rankings.append([c.chunk_id for c in chunks])
            weights.append(q_weight)
            for c in chunks:
                old = chunk_map.get(c.chunk_id)
                if old is None:
                    chunk_map[c.chunk_id] = c
                else:
                    if c.distance < old.distance:
                        old.distance = c.distance
                    if c.semantic_score > old.semantic_score:
                        old.semantic_score = c.semantic_score
            use_bm25 = (self.enable_hybrid and self.chunk_bm25 is not None and self.chunk_bm25.is_built
                        and (bm25_all_queries or q_idx == 0))
            if use_bm25:
                hits = self.chunk_bm25.search(q, top_k=n_fetch, doc_filter=doc_filter)
                rankings.append([h.doc_key for h in hits])
                if bm25_all_queries:
                    weights.append(q_weight * 1.5 if q_idx == 0 else q_weight * 0.75)
                else:
                    weights.append(q_weight)
                for h in hits:
                    if h.doc_key not in chunk_map:
                        meta = dict(h.metadata)
                        text = meta.pop("text", "")
                        c = _chunk_from_meta(h.doc_key, text, meta, 1.0)
                        chunk_map[h.doc_key] = c
                    c = chunk_map[h.doc_key]
                    c.bm25_score = max(c.bm25_score, h.score) if bm25_all_queries else h.score
        if len(rankings) > 1:
            fused = self.fuse(rankings, weights=weights)
            for cid, c in chunk_map.items():
                c.hybrid_score = fused.get(cid, 0.0)
        else:
            for c in chunk_map.values():
                c.hybrid_score = c.semantic_score
        ordered = list(chunk_map.values())
        ordered.sort(key=lambda c: c.hybrid_score, reverse=True)
        return ordered

    def retrieve_candidates(self, query: str, n_candidates: int = 100,
                            where_filter: Optional[Dict[str, Any]] = None) -> List[RetrievedChunk]:
        ordered = self._gather(query, where_filter, n_fetch=max(n_candidates, 50), backfill=10,
                               bm25_all_queries=True)
        return ordered[:n_candidates]

    def retrieve(self, query: str, where_filter: Optional[Dict[str, Any]] = None,
                 n_documents: Optional[int] = None, n_chunks_per_doc: Optional[int] = None) -> List[RetrievedDocument]:
        n_docs = n_documents or self.n_documents
        n_chunks = n_chunks_per_doc or self.n_chunks_per_doc
        ordered = self._gather(query, where_filter, n_fetch=n_docs * self.fetch_multiplier, backfill=5,
                               bm25_all_queries=False)
        return self._deduplicate_by_document(ordered, n_docs, n_chunks)

    def _deduplicate_by_document(self, chunks, n_documents, n_chunks_per_doc):
        by_doc = defaultdict(list)
        for c in chunks:
            by_doc[c.document_path].append(c)
        docs, seen_urls = [], set()
        for path, group in by_doc.items():
            best = sorted(group, key=lambda c: c.hybrid_score if c.hybrid_score > 0 else c.similarity_score,
                          reverse=True)[:n_chunks_per_doc]
            url = best[0].metadata.get("source_url", "") if best else ""
            if url:
                norm = url.lower().replace("https://", "").replace("http://", "").replace("www.", "")
                if norm in seen_urls:
                    continue
                seen_urls.add(norm)
            docs.append(RetrievedDocument(document_path=path, chunks=best, avg_similarity=0.0, primary_nature=""))
        docs.sort(key=lambda d: d.avg_similarity, reverse=True)
        return docs[:n_documents]


# --------------------------------------------------------------------------------------------------
# Batched front-end.  The reference issues, per question, 4 dense + 4 BM25 calls
# one query at a time (src/rag/retriever.py:372-452).  retrieve_candidates_batch answers many questions
# with ONE dense call (all query variants of all questions), ONE BM25 call per distinct doc_filter and
# ONE device RRF call, and only materialises RetrievedChunk objects for the final candidates.  Results
# are identical to calling retrieve_candidates() question by question.
# --------------------------------------------------------------------------------------------------
def _retrieve_candidates_batch(self, queries, n_candidates: int = 100, where_filter=None):
    import numpy as np
    from .collection import distance_from_score
    from .rrf import rrf_fuse_rows

    col, bm = self.collection, self.chunk_bm25
    use_bm25 = self.enable_hybrid and bm is not None and bm.is_built
    if not hasattr(col, "query_rows") or (use_bm25 and not hasattr(bm, "search_rows")):
        return [self.retrieve_candidates(q, n_candidates, where_filter) for q in queries]
    n_fetch = max(n_candidates, 50)
    from .rrf import rrf_capacity
    # 1. query variants + summary pre-filter per question
    variants, filters = [], []
    for q in queries:
        expanded = self.acronym_expander(q)
        variants.append(self.query_expander.expand(expanded) if self.query_expander is not None else [expanded])
        f = None
        if self.enable_summary_prefilter and self.summary_bm25 is not None and self.summary_bm25._is_built:
            f = self.summary_bm25.get_relevant_doc_paths(expanded, top_k=self.summary_prefilter_k)
        filters.append(f)
    flat = [v for vs in variants for v in vs]
    owner = [qi for qi, vs in enumerate(variants) for _ in vs]
    # 2. one dense call for every variant of every question
    emb = np.asarray(self._embed(flat), dtype=np.float32)
    d_rows, d_scores, d_counts = col.query_rows(emb, n_fetch, where_filter)
    # 3. BM25: one call per distinct doc_filter (the filter applies BEFORE the top-k, bm25_index.py:272-275)
    b_rows = b_scores = b_counts = None
    if use_bm25:
        k_b = min(n_fetch, len(bm.chunk_ids))
        tokens = [bm.tokenizer(v) for v in flat]
        b_rows = np.full((len(flat), max(k_b, 1)), -1, np.int32)
        b_scores = np.zeros((len(flat), max(k_b, 1)), np.float64)
        b_counts = np.zeros(len(flat), np.int32)
        groups = {}
        for vi, qi in enumerate(owner):
            key = frozenset(filters[qi]) if filters[qi] is not None else None
            groups.setdefault(key, []).append(vi)
        for key, idx in groups.items():
            live = [vi for vi in idx if tokens[vi]]
            if not live or k_b == 0:
                continue
            r, s, c = bm.search_rows([tokens[vi] for vi in live], n_fetch, set(key) if key is not None else None)
            b_rows[live, :r.shape[1]], b_scores[live, :r.shape[1]], b_counts[live] = r, s, c
        # BM25 row -> collection row (-1: the chunk left the collection since the index was built).  Rebuilt
        # whenever the collection mutates or the index is rebuilt: rows shift on delete/compaction.
        stamp = (id(bm.index), getattr(col, "mutation_version", None))
        if getattr(bm, "_to_col_stamp", None) != stamp or stamp[1] is None:
            pos = col._pos
            bm._to_col = np.array([pos.get(cid, -1) for cid in bm.chunk_ids], dtype=np.int32)
            bm._to_col_stamp = stamp
        if len(bm._to_col) and int(bm._to_col.min()) < 0:
            # a stale keyword index (chunks deleted from the collection since it was built): the per-question path
            # fuses by chunk id and still reports those hits from the index's own metadata — take it
            return [self.retrieve_candidates(q, n_candidates, where_filter) for q in queries]
    # 4. rankings in the reference's order: dense v0, bm25 v0, dense v1, bm25 v1, ...
    Q = len(queries)
    vmax = max(len(vs) for vs in variants)
    R = vmax * (2 if use_bm25 else 1)
    if R * n_fetch > rrf_capacity():        # more ranking entries than one device RRF call holds
        return [self.retrieve_candidates(q, n_candidates, where_filter) for q in queries]
    ids = np.full((Q, R, n_fetch), -1, np.int32)
    weights = np.zeros((Q, R), np.float64)
    dense_lists = {}
    paths = None
    vi = 0
    for qi in range(Q):
        for v in range(len(variants[qi])):
            rows = d_rows[vi, :d_counts[vi]].tolist()
            if filters[qi]:
                if paths is None:
                    paths = [(m or {}).get("document_path", "") for m in col._metas]
                kept = [r for r in rows if paths[r] in filters[qi]]
                if len(kept) < 10:
                    ks = set(kept)
                    kept.extend([r for r in rows if r not in ks][:10 - len(kept)])
                rows = kept
            dense_lists[(qi, v)] = (rows, vi)
            slot = v * 2 if use_bm25 else v
            ids[qi, slot, :len(rows)] = rows
            q_weight = 2.0 if v == 0 else 1.0
            weights[qi, slot] = q_weight
            if use_bm25:
                ids[qi, slot + 1, :b_counts[vi]] = bm._to_col[b_rows[vi, :b_counts[vi]]]
                weights[qi, slot + 1] = q_weight * 1.5 if v == 0 else q_weight * 0.75
            vi += 1
    # 5. one device RRF call for all questions
    f_ids, f_scores, f_counts = rrf_fuse_rows(ids, weights, 60, n_candidates)
    # 6. materialise only the final candidates
    out = []
    dist_all = (1.0 - d_scores).astype(np.float32)                # distance_from_score, vectorised
    for qi in range(Q):
        nv = len(variants[qi])
        best_dist, best_bm25 = {}, {}
        for v in range(nv):
            rows, vi = dense_lists[(qi, v)]
            all_rows = d_rows[vi, :d_counts[vi]].tolist()
            dists = dist_all[vi, :d_counts[vi]].tolist()          # float32(1 - score) as python floats
            if len(rows) == len(all_rows):                        # no doc_filter: the list as returned
                pairs = zip(all_rows, dists)
            else:
                pos = dict(zip(all_rows, dists))
                pairs = ((r, pos[r]) for r in rows)
            for r, dist in pairs:
                if r not in best_dist or dist < best_dist[r]:
                    best_dist[r] = dist
            if use_bm25:
                for j in range(b_counts[vi]):
                    r = int(bm._to_col[b_rows[vi, j]])
                    best_bm25[r] = max(best_bm25.get(r, 0.0), float(b_scores[vi, j]))
        single = (nv * (2 if use_bm25 else 1)) <= 1
        chunks = []
        for j in range(f_counts[qi]):
            r = int(f_ids[qi, j])
            meta = col._metas[r] or {}
            dist = best_dist.get(r, 1.0)
            c = _chunk_from_meta(col._ids[r], col._docs[r], meta if r in best_dist else dict(meta), dist)
            c.semantic_score = c.similarity_score if r in best_dist else 0.0
            c.bm25_score = best_bm25.get(r, 0.0)
            c.hybrid_score = c.semantic_score if single else float(f_scores[qi, j])
            chunks.append(c)
        if single:
            chunks.sort(key=lambda c: c.hybrid_score, reverse=True)
        out.append(chunks)
    return out


HybridRetriever.retrieve_candidates_batch = _retrieve_candidates_batch
