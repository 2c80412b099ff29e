"""shared test helpers (checker side: may import oracle/)."""
import numpy as np

from oracle import numpy_oracle as no


class FixedEmbeddingProvider:
    def __init__(self, table):
        self.table = table

    def embed(self, texts):
        return [list(map(float, self.table[t])) for t in texts]


class FixedQueryExpander:
    def __init__(self, table):
        self.table = table

    def expand(self, query):
        return [query] + list(self.table.get(query, []))


def acronym_expander_for_golden(golden):
    """Acronym expansion (src/utils/acronyms.py:151-198) is host string prep upstream
    of the boundary; the golden records what the reference's own function returned
    for each fixture query."""
    table = golden["expanded"]
    return lambda q: table.get(q, q)


def fill(col, chunks, emb, batch=100):
    for s in range(0, len(chunks), batch):
        part = chunks[s:s + batch]
        col.add(ids=[c["id"] for c in part], documents=[c["text"] for c in part],
                embeddings=emb[s:s + batch], metadatas=[c["metadata"] for c in part])


def chunk_dump(c):
    return {"chunk_id": c.chunk_id, "distance": float(c.distance).hex(), "semantic": float(c.semantic_score).hex(),
            "bm25": float(c.bm25_score).hex(), "hybrid": float(c.hybrid_score).hex(),
            "document_path": c.document_path, "chunk_index": c.chunk_index}


def synth_unit(n, d, seed):
    g = np.random.default_rng(seed)
    return no.l2_normalize_rows(g.standard_normal((n, d)).astype(np.float32))


def zipf_docs(n_docs, vocab, seed, lo=40, hi=250, s=1.07):
    """synthetic tokenised corpus: Zipf(s) term ids, relabelled in first-seen order."""
    g = np.random.default_rng(seed)
    lens = g.integers(lo, hi + 1, size=n_docs)
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    p = ranks ** (-s)
    p /= p.sum()
    flat = g.choice(vocab, size=int(lens.sum()), p=p)
    # relabel in first-seen order (rank-bm25's idf sum order)
    _, first = np.unique(flat, return_index=True)
    order = np.argsort(first)
    remap = np.empty(vocab, dtype=np.int64)
    remap[:] = -1
    seen_terms = np.unique(flat)[order]
    remap[seen_terms] = np.arange(len(seen_terms))
    flat = remap[flat]
    docs = np.split(flat, np.cumsum(lens)[:-1])
    return docs, len(seen_terms)


class OracleChunkBM25Index:
    """CPU stand-in with ChunkBM25Index semantics (src/rag/bm25_index.py:176-296) built on the
    restated rank_bm25 — the checker for DeviceChunkBM25Index and the CPU leg of host-logic tests."""

    def __init__(self, tokenizer):
        from oracle import rank_bm25 as rb
        self._rb = rb
        self.tokenizer = tokenizer
        self._is_built = False

    @property
    def is_built(self):
        return self._is_built

    def build_from_collection(self, collection, batch_size=5000):
        from b200rag.bm25 import BM25Result
        self._Result = BM25Result
        self.chunk_ids, self.chunk_texts, self.chunk_metadatas, self.corpus_tokens = [], [], [], []
        total, offset = collection.count(), 0
        while offset < total:
            batch = collection.get(limit=batch_size, offset=offset, include=["documents", "metadatas"])
            for cid, text, meta in zip(batch["ids"], batch["documents"], batch["metadatas"]):
                if not text or not text.strip():
                    continue
                toks = self.tokenizer(text)
                if not toks:
                    continue
                self.chunk_ids.append(cid); self.chunk_texts.append(text)
                self.chunk_metadatas.append(meta); self.corpus_tokens.append(toks)
            offset += batch_size
        self.index = self._rb.BM25Okapi(self.corpus_tokens)
        self._is_built = True

    def search(self, query, top_k=30, doc_filter=None):
        toks = self.tokenizer(query)
        if not toks:
            return []
        scores = self.index.get_scores(toks)
        scored = [(i, scores[i]) for i in range(len(scores)) if scores[i] > 0 and
                  (doc_filter is None or self.chunk_metadatas[i].get("document_path", "") in doc_filter)]
        scored.sort(key=lambda t: t[1], reverse=True)
        return [self._Result(doc_key=self.chunk_ids[i], score=float(s),
                             metadata={**self.chunk_metadatas[i], "text": self.chunk_texts[i]})
                for i, s in scored[:top_k]]


class OracleSummaryBM25Index:
    def __init__(self, tokenizer):
        from oracle import rank_bm25 as rb
        self._rb, self.tokenizer, self._is_built = rb, tokenizer, False

    def build(self, path):
        import json
        with open(path, "r", encoding="utf-8") as f:
            summaries = json.load(f)
        self.doc_keys, toks = [], []
        for p, e in summaries.items():
            s = e.get("summary", "")
            if not s or s.startswith("ERREUR"):
                continue
            t = self.tokenizer(f"{e.get('document_title', '')} {s} {e.get('source_url', '')}")
            if not t:
                continue
            self.doc_keys.append(p); toks.append(t)
        self.index = self._rb.BM25Okapi(toks)
        self._is_built = True

    def search_pairs(self, query, top_k):
        toks = self.tokenizer(query)
        if not toks:
            return []
        scores = self.index.get_scores(toks)
        scored = [(i, scores[i]) for i in range(len(scores)) if scores[i] > 0]
        scored.sort(key=lambda t: t[1], reverse=True)
        return [(self.doc_keys[i], float(s)) for i, s in scored[:top_k]]

    def get_relevant_doc_paths(self, query, top_k=20):
        return {k for k, _ in self.search_pairs(query, top_k)}


def oracle_fuse(rankings, k=60, weights=None):
    from oracle import numpy_oracle as no
    ids, sc = no.rrf_fuse(rankings, weights, k)
    return dict(zip(ids, sc))


def run_e2e_case(retriever_cls, col, bm, sm, gold, table, run, fuse=None):
    cfg = run["config"]
    kw = {} if fuse is None else {"fuse": fuse}
    r = retriever_cls(collection=col, llm_provider=None, embedding_provider=FixedEmbeddingProvider(table),
                      summary_bm25_index=sm if cfg["prefilter"] else None, chunk_bm25_index=bm,
                      query_expander=FixedQueryExpander(gold["expansions"]) if cfg["expander"] else None,
                      n_documents=5, n_chunks_per_doc=3, summary_prefilter_k=8, enable_hybrid=cfg["hybrid"],
                      enable_summary_prefilter=cfg["prefilter"],
                      acronym_expander=acronym_expander_for_golden(gold), **kw)
    cands = r.retrieve_candidates(run["query"], n_candidates=40, where_filter=run["where"])
    docs = r.retrieve(run["query"], where_filter=run["where"])
    got_docs = [{"document_path": d.document_path, "avg_similarity": float(d.avg_similarity).hex(),
                 "primary_nature": d.primary_nature, "chunks": [chunk_dump(c) for c in d.chunks]} for d in docs]
    for g, want in zip(got_docs, run["documents"]):
        if want["primary_nature"] is None:      # tie broken by str hash order in the reference: unpinned
            g["primary_nature"] = None
    return [chunk_dump(c) for c in cands], got_docs
