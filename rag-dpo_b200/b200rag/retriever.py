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
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

from . import rrf as _rrf


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
            except Exception:
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
