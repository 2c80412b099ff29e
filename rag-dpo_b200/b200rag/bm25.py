"""Device BM25 indexes — drop-ins for ChunkBM25Index / SummaryBM25Index
(src/rag/bm25_index.py:60-173, 176-296).

Host side (this file): tokenisation, vocabulary (first-seen order), CSR
postings, the rank-bm25 0.2.2 idf table (natural log, +0.5 terms, negative
idfs floored to epsilon * average_idf — bm25_index.py:126,236 construct
BM25Okapi with its defaults k1=1.5, b=0.75, epsilon=0.25).
Device side (csrc/bm25.cu): scoring + select, fp64, bit-identical to numpy.
"""
import ctypes as C
import json
import math
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Set

import numpy as np

from . import _lib
from .tokenizer import tokenize_french
from .where import bitmap_from_mask


@dataclass
class BM25Result:
    """same fields as the reference's BM25Result (src/rag/bm25_index.py:52-57)"""
    doc_key: str
    score: float
    metadata: Dict


def _pack_queries(queries_term_ids):
    """list of term-id sequences -> (flat int32 terms, q_ptr int32[Q + 1]); one concatenate over the list (the
    per-query conversions it replaces were half of the host time of a 256-query call)"""
    Q = len(queries_term_ids)
    q_ptr = np.zeros(Q + 1, dtype=np.int32)
    np.cumsum(np.fromiter(map(len, queries_term_ids), dtype=np.int64, count=Q), out=q_ptr[1:])
    if not q_ptr[-1]:
        return np.zeros(1, np.int32), q_ptr
    flat = np.concatenate(queries_term_ids) if Q > 1 else np.asarray(queries_term_ids[0])
    return np.ascontiguousarray(flat, dtype=np.int32), q_ptr


class Postings:
    """CSR postings + BM25Okapi statistics built on the host."""

    def __init__(self, term_ptr, post_row, post_tf, doc_len, idf, avgdl, k1, b, epsilon, vocab=None):
        self.term_ptr = np.ascontiguousarray(term_ptr, dtype=np.int64)
        self.post_row = np.ascontiguousarray(post_row, dtype=np.int32)
        self.post_tf = np.ascontiguousarray(post_tf, dtype=np.int32)
        self.doc_len = np.ascontiguousarray(doc_len, dtype=np.int32)
        self.idf = np.ascontiguousarray(idf, dtype=np.float64)
        self.avgdl, self.k1, self.b, self.epsilon = float(avgdl), float(k1), float(b), float(epsilon)
        self.vocab = vocab

    @property
    def n_docs(self):
        return len(self.doc_len)

    @property
    def n_terms(self):
        return len(self.idf)

    @staticmethod
    def idf_table(df, n_docs, epsilon):
        """rank-bm25 0.2.2 _calc_idf: python floats, math.log, sum in vocabulary order."""
        idf = [0.0] * len(df)
        total = 0
        present = 0
        negative = []
        for t, n_t in enumerate(df):
            n_t = int(n_t)
            if n_t == 0:
                continue
            v = math.log(n_docs - n_t + 0.5) - math.log(n_t + 0.5)
            idf[t] = v
            total += v
            present += 1
            if v < 0:
                negative.append(t)
        if present:
            floor = epsilon * (total / present)
            for t in negative:
                idf[t] = floor
        return np.array(idf, dtype=np.float64)

    @classmethod
    def from_flat_tokens(cls, doc_ptr, tokens, n_terms=None, k1=1.5, b=0.75, epsilon=0.25, vocab=None, threads=0):
        """documents as ONE int32 array of term ids (vocabulary = first-seen order) + n_docs+1 offsets.  The CSR
        is built by the library's multi-threaded host code (rag_csr_build, csrc/host_csr.cu; no GPU involved)."""
        doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        n_docs = len(doc_ptr) - 1
        if n_docs <= 0:
            raise ValueError("cannot build a BM25 index over zero documents")
        if n_terms is None:
            n_terms = int(tokens.max()) + 1 if len(tokens) else 0
        total = int(doc_ptr[-1])
        term_ptr = np.zeros(n_terms + 1, dtype=np.int64)
        post_row = np.empty(max(total, 1), dtype=np.int32)
        post_tf = np.empty(max(total, 1), dtype=np.int32)
        nnz = C.c_int64()
        rc = _lib.load().rag_csr_build(n_docs, _lib.ptr(doc_ptr), _lib.ptr(tokens) if total else None, n_terms,
                                       _lib.ptr(term_ptr), _lib.ptr(post_row), _lib.ptr(post_tf), len(post_row),
                                       C.byref(nnz), int(threads))
        if rc != 0:
            raise ValueError(f"rag_csr_build failed ({rc}): term ids must lie in [0, n_terms)")
        lens = np.diff(doc_ptr)
        df = np.diff(term_ptr)
        avgdl = int(lens.sum()) / n_docs
        idf = cls.idf_table(df, n_docs, epsilon)
        return cls(term_ptr, post_row[:nnz.value].copy(), post_tf[:nnz.value].copy(), lens, idf, avgdl, k1, b, epsilon,
                   vocab)

    @classmethod
    def from_term_ids(cls, docs_term_ids, n_terms=None, k1=1.5, b=0.75, epsilon=0.25, vocab=None):
        """docs_term_ids: sequence of 1-D int arrays, ids in vocabulary (first-seen) order."""
        n_docs = len(docs_term_ids)
        if n_docs == 0:
            raise ValueError("cannot build a BM25 index over zero documents")
        lens = np.fromiter((len(d) for d in docs_term_ids), dtype=np.int64, count=n_docs)
        doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
        np.cumsum(lens, out=doc_ptr[1:])
        flat = (np.concatenate([np.asarray(d, dtype=np.int32) for d in docs_term_ids])
                if doc_ptr[-1] else np.zeros(0, np.int32))
        return cls.from_flat_tokens(doc_ptr, flat, n_terms=n_terms, k1=k1, b=b, epsilon=epsilon, vocab=vocab)

    @classmethod
    def from_token_lists(cls, corpus_tokens, **kw):
        """corpus_tokens: list of token lists (what ChunkBM25Index keeps, src/rag/bm25_index.py:186); the
        vocabulary is numbered in first-seen order (the order rank-bm25 sums the idfs in)."""
        vocab = {}
        n_docs = len(corpus_tokens)
        doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
        np.cumsum([len(t) for t in corpus_tokens], out=doc_ptr[1:])
        flat = np.empty(int(doc_ptr[-1]), dtype=np.int32)
        setdefault = vocab.setdefault
        pos = 0
        for toks in corpus_tokens:
            for w in toks:
                flat[pos] = setdefault(w, len(vocab))
                pos += 1
        return cls.from_flat_tokens(doc_ptr, flat, n_terms=len(vocab), vocab=vocab, **kw)

    def term_ids(self, tokens):
        return np.array([self.vocab.get(w, -1) for w in tokens], dtype=np.int32)


class DeviceBM25:
    """Owner of a rag_bm25_t handle.  n_shards > 1: the documents are spread over the GPUs of this process like a
    sharded corpus (every GPU holds the postings of its rows; idf / avgdl are the global statistics)."""

    def __init__(self, postings: Postings, n_shards=1, devices=None):
        self.p = postings
        self.n_shards = int(n_shards)
        self._L = _lib.lib()
        h = C.c_void_p()
        p = postings
        if self.n_shards > 1:
            _lib.ensure_slots(self.n_shards, devices)
            _lib.check(self._L.rag_bm25_create_sharded(C.byref(h), self.n_shards, p.n_docs, p.n_terms, len(p.post_row),
                                                       _lib.ptr(p.term_ptr), _lib.ptr(p.post_row), _lib.ptr(p.post_tf),
                                                       _lib.ptr(p.doc_len), _lib.ptr(p.idf), p.avgdl, p.k1, p.b))
        else:
            _lib.check(self._L.rag_bm25_create(C.byref(h), p.n_docs, p.n_terms, len(p.post_row), _lib.ptr(p.term_ptr),
                                               _lib.ptr(p.post_row), _lib.ptr(p.post_tf), _lib.ptr(p.doc_len),
                                               _lib.ptr(p.idf), p.avgdl, p.k1, p.b))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.rag_bm25_destroy(self._h)
            self._h = None

    __del__ = close

    def search_ids(self, queries_term_ids, k, allow_bitmap=None):
        """queries_term_ids: list of int32 arrays.  Returns rows (Q,k) int32 [-1 padded],
        scores (Q,k) fp64, counts (Q,) int32."""
        Q = len(queries_term_ids)
        if k > _lib.RAG_MAX_K:
            return self._search_multipass(queries_term_ids, int(k), allow_bitmap)
        flat, q_ptr = _pack_queries(queries_term_ids)
        rows = np.empty((Q, k), dtype=np.int32)
        scores = np.empty((Q, k), dtype=np.float64)
        counts = np.empty(Q, dtype=np.int32)
        ab = np.ascontiguousarray(allow_bitmap, dtype=np.uint8) if allow_bitmap is not None else None
        _lib.check(self._L.rag_bm25_search(self._h, _lib.ptr(flat), _lib.ptr(q_ptr), Q, int(k), _lib.ptr(ab),
                                           _lib.ptr(rows), _lib.ptr(scores), _lib.ptr(counts)))
        return rows, scores, counts

    def query_bytes(self, queries_term_ids):
        """measurement helper: (bytes the filter pass streams, postings) per query, int64 arrays"""
        Q = len(queries_term_ids)
        flat, q_ptr = _pack_queries(queries_term_ids)
        nbytes = np.zeros(Q, dtype=np.int64)
        npost = np.zeros(Q, dtype=np.int64)
        _lib.check(self._L.rag_bm25_query_bytes(self._h, _lib.ptr(flat), _lib.ptr(q_ptr), Q, _lib.ptr(nbytes), _lib.ptr(npost)))
        return nbytes, npost

    def _search_multipass(self, queries_term_ids, k, allow_bitmap):
        """top_k above RAG_MAX_K: successive passes with the rows already returned masked out (still exact)"""
        n, Q = self.p.n_docs, len(queries_term_ids)
        rows = np.full((Q, k), -1, dtype=np.int32)
        scores = np.zeros((Q, k), dtype=np.float64)
        counts = np.zeros(Q, dtype=np.int32)
        base = (np.unpackbits(np.asarray(allow_bitmap, dtype=np.uint8), bitorder="little")[:n].astype(bool)
                if allow_bitmap is not None else np.ones(n, dtype=bool))
        for qi, terms in enumerate(queries_term_ids):
            mask, got = base.copy(), 0
            while got < k:
                kk = min(_lib.RAG_MAX_K, k - got)
                r, s, c = self.search_ids([terms], kk, np.packbits(mask, bitorder="little"))
                c0 = int(c[0])
                rows[qi, got:got + c0], scores[qi, got:got + c0] = r[0, :c0], s[0, :c0]
                mask[r[0, :c0]] = False
                got += c0
                if c0 < kk:
                    break
            counts[qi] = got
        return rows, scores, counts

    def bytes_per_posting(self):
        """bytes the search kernels stream per posting (the algorithmic bytes of the roofline): 4 on the packed
        filter path, 12 (row + fp64 impact) when only the exact range path serves this index"""
        b = C.c_int()
        _lib.check(self._L.rag_bm25_info(self._h, C.byref(b), None))
        return b.value

    def index_bytes(self):
        n = C.c_int64()
        _lib.check(self._L.rag_bm25_info(self._h, None, C.byref(n)))
        return n.value

    def scores(self, term_ids):
        """full fp64 score vector (BM25Okapi.get_scores)"""
        t = np.ascontiguousarray(term_ids, dtype=np.int32)
        out = np.empty(self.p.n_docs, dtype=np.float64)
        _lib.check(self._L.rag_bm25_scores(self._h, _lib.ptr(t) if len(t) else None, len(t), _lib.ptr(out)))
        return out


def _clamp_k(top_k, n):
    return min(int(top_k), n)        # top_k above RAG_MAX_K is served in several passes (DeviceBM25.search_ids)


class DeviceChunkBM25Index:
    """Drop-in for ChunkBM25Index (src/rag/bm25_index.py:176-296)."""

    def __init__(self, tokenizer=tokenize_french, n_shards=1, devices=None):
        self.tokenizer = tokenizer
        self.n_shards, self.devices = int(n_shards), devices
        self.index: Optional[DeviceBM25] = None
        self.chunk_ids: List[str] = []
        self.chunk_texts: List[str] = []
        self.chunk_metadatas: List[Dict] = []
        self.corpus_tokens: List[List[str]] = []
        self._is_built = False
        self._filter_cache = {}

    def build_from_collection(self, collection, batch_size: int = 5000) -> None:
        total = collection.count()
        self.chunk_ids, self.chunk_texts, self.chunk_metadatas, self.corpus_tokens = [], [], [], []
        offset = 0
        while offset < total:
            batch = collection.get(limit=batch_size, offset=offset, include=["documents", "metadatas"])
            for chunk_id, text, metadata in zip(batch["ids"], batch["documents"], batch["metadatas"]):
                if not text or not text.strip():
                    continue
                tokens = self.tokenizer(text)
                if not tokens:
                    continue
                self.chunk_ids.append(chunk_id)
                self.chunk_texts.append(text)
                self.chunk_metadatas.append(metadata)
                self.corpus_tokens.append(tokens)
            offset += batch_size
        self._vocab, self._flat, self._doc_ptr = {}, np.zeros(0, np.int32), np.zeros(1, np.int64)
        self._append_tokens(self.corpus_tokens)
        self._finish(Postings.from_flat_tokens(self._doc_ptr, self._flat, n_terms=len(self._vocab), vocab=self._vocab))

    # ---- incremental maintenance (the reference rebuilds from the whole collection, src/rag/pipeline.py:1031-1036;
    # these keep the tokens of the chunks already indexed and only tokenise what is new) ------------------------
    def _append_tokens(self, token_lists):
        """term ids of new documents appended to the cached flat array; the vocabulary keeps its first-seen order"""
        n_new = sum(len(t) for t in token_lists)
        flat = np.empty(n_new, dtype=np.int32)
        setdefault, vocab = self._vocab.setdefault, self._vocab
        pos = 0
        for toks in token_lists:
            for w in toks:
                flat[pos] = setdefault(w, len(vocab))
                pos += 1
        ptr = np.zeros(len(token_lists), dtype=np.int64)
        if len(token_lists):
            np.cumsum([len(t) for t in token_lists], out=ptr)
        self._doc_ptr = np.concatenate([self._doc_ptr, self._doc_ptr[-1] + ptr])
        self._flat = np.concatenate([self._flat, flat])

    def add_chunks(self, ids, documents, metadatas=None) -> int:
        """collection.add for the keyword index (enterprise ingestion, src/processing/ingest_enterprise.py:241-246):
        the new chunks are tokenised, the CSR is rebuilt natively from the cached term ids (rag_csr_build) and the
        device index replaced.  Identical to build_from_collection over the grown collection.  Returns the number
        of chunks indexed (empty texts are skipped, as in the build)."""
        if not self._is_built or not hasattr(self, "_vocab"):
            raise RuntimeError("Index non construit. Appelez build_from_collection() d'abord.")
        metadatas = metadatas if metadatas is not None else [{} for _ in ids]
        new_tokens = []
        for chunk_id, text, metadata in zip(ids, documents, metadatas):
            if not text or not text.strip():
                continue
            tokens = self.tokenizer(text)
            if not tokens:
                continue
            self.chunk_ids.append(chunk_id)
            self.chunk_texts.append(text)
            self.chunk_metadatas.append(metadata)
            self.corpus_tokens.append(tokens)
            new_tokens.append(tokens)
        if new_tokens:
            self._append_tokens(new_tokens)
            self._finish(Postings.from_flat_tokens(self._doc_ptr, self._flat, n_terms=len(self._vocab), vocab=self._vocab))
        return len(new_tokens)

    def remove_chunks(self, ids) -> int:
        """collection.delete for the keyword index (src/processing/ingest_enterprise.py:272,304): the chunks leave
        the index; the vocabulary is renumbered in first-seen order of what remains (what a fresh build yields),
        nothing is re-tokenised."""
        if not self._is_built or not hasattr(self, "_vocab"):
            raise RuntimeError("Index non construit. Appelez build_from_collection() d'abord.")
        gone = set(ids)
        keep = [i for i, cid in enumerate(self.chunk_ids) if cid not in gone]
        removed = len(self.chunk_ids) - len(keep)
        if removed == 0:
            return 0
        if not keep:
            raise ValueError("cannot remove every chunk of a BM25 index")
        self.chunk_ids = [self.chunk_ids[i] for i in keep]
        self.chunk_texts = [self.chunk_texts[i] for i in keep]
        self.chunk_metadatas = [self.chunk_metadatas[i] for i in keep]
        self.corpus_tokens = [self.corpus_tokens[i] for i in keep]
        self._vocab, self._flat, self._doc_ptr = {}, np.zeros(0, np.int32), np.zeros(1, np.int64)
        self._append_tokens(self.corpus_tokens)
        self._finish(Postings.from_flat_tokens(self._doc_ptr, self._flat, n_terms=len(self._vocab), vocab=self._vocab))
        return removed

    def build_from_postings(self, postings, chunk_ids=None, chunk_texts=None, chunk_metadatas=None):
        """Index pre-tokenised / synthetic corpora (integer term ids)."""
        n = postings.n_docs
        self.chunk_ids = list(chunk_ids) if chunk_ids is not None else [str(i) for i in range(n)]
        self.chunk_texts = list(chunk_texts) if chunk_texts is not None else [""] * n
        self.chunk_metadatas = list(chunk_metadatas) if chunk_metadatas is not None else [{} for _ in range(n)]
        self.corpus_tokens = []
        self._finish(postings)

    def _finish(self, postings):
        if self.index is not None:
            self.index.close()
        self.postings = postings
        self.index = DeviceBM25(postings, n_shards=self.n_shards, devices=self.devices)
        self._filter_cache = {}
        self._to_col_stamp = None          # the batched front-end's row map belongs to the previous index
        self._is_built = True

    @property
    def is_built(self) -> bool:
        return self._is_built

    def _doc_filter_bitmap(self, doc_filter):
        key = frozenset(doc_filter)
        hit = self._filter_cache.get(key)
        if hit is None:
            mask = np.fromiter(((m or {}).get("document_path", "") in key for m in self.chunk_metadatas),
                               dtype=bool, count=len(self.chunk_metadatas))
            hit = bitmap_from_mask(mask)
            if len(self._filter_cache) > 32:
                self._filter_cache.clear()
            self._filter_cache[key] = hit
        return hit

    def search_rows(self, query_tokens_list, top_k, doc_filter=None):
        k = _clamp_k(top_k, len(self.chunk_ids))
        ids = [self.postings.term_ids(t) for t in query_tokens_list]
        bitmap = self._doc_filter_bitmap(doc_filter) if doc_filter is not None else None
        return self.index.search_ids(ids, k, bitmap)

    def search(self, query: str, top_k: int = 30, doc_filter: Optional[Set[str]] = None) -> List[BM25Result]:
        if not self._is_built:
            raise RuntimeError("Index non construit. Appelez build_from_collection() d'abord.")
        tokens = self.tokenizer(query)
        if not tokens or top_k <= 0:
            return []
        rows, scores, counts = self.search_rows([tokens], top_k, doc_filter)
        out = []
        for r, s in zip(rows[0, :counts[0]].tolist(), scores[0, :counts[0]].tolist()):
            out.append(BM25Result(doc_key=self.chunk_ids[r], score=s,
                                  metadata={**self.chunk_metadatas[r], "text": self.chunk_texts[r]}))
        return out


class DeviceSummaryBM25Index:
    """Drop-in for SummaryBM25Index (src/rag/bm25_index.py:60-173)."""

    def __init__(self, summaries_path: Optional[Path] = None, tokenizer=tokenize_french):
        self.summaries_path = Path(summaries_path) if summaries_path else Path("data/keep/cnil/document_summaries.json")
        self.tokenizer = tokenizer
        self.index: Optional[DeviceBM25] = None
        self.doc_keys: List[str] = []
        self.doc_metadata: List[Dict] = []
        self.corpus_tokens: List[List[str]] = []
        self._is_built = False

    def build(self, summaries_path: Optional[str] = None) -> None:
        if summaries_path:
            self.summaries_path = Path(summaries_path)
        if not self.summaries_path.exists():
            raise FileNotFoundError(f"Fichier summaries introuvable: {self.summaries_path}")
        with open(self.summaries_path, "r", encoding="utf-8") as f:
            summaries = json.load(f)
        self.doc_keys, self.doc_metadata, self.corpus_tokens = [], [], []
        for doc_path, entry in summaries.items():
            summary_text = entry.get("summary", "")
            if not summary_text or summary_text.startswith("ERREUR"):
                continue
            title = entry.get("document_title", "")
            url = entry.get("source_url", "")
            tokens = self.tokenizer(f"{title} {summary_text} {url}")
            if not tokens:
                continue
            self.doc_keys.append(doc_path)
            self.doc_metadata.append({"document_path": doc_path, "source_url": url, "document_title": title,
                                      "summary": summary_text})
            self.corpus_tokens.append(tokens)
        self.postings = Postings.from_token_lists(self.corpus_tokens)
        if self.index is not None:
            self.index.close()
        self.index = DeviceBM25(self.postings)
        self._is_built = True

    def search(self, query: str, top_k: int = 20) -> List[BM25Result]:
        if not self._is_built:
            raise RuntimeError("Index non construit. Appelez build() d'abord.")
        tokens = self.tokenizer(query)
        if not tokens or top_k <= 0:
            return []
        k = _clamp_k(top_k, len(self.doc_keys))
        rows, scores, counts = self.index.search_ids([self.postings.term_ids(tokens)], k)
        return [BM25Result(doc_key=self.doc_keys[r], score=s, metadata=self.doc_metadata[r])
                for r, s in zip(rows[0, :counts[0]].tolist(), scores[0, :counts[0]].tolist())]

    def get_relevant_doc_paths(self, query: str, top_k: int = 20) -> Set[str]:
        return {r.doc_key for r in self.search(query, top_k=top_k)}
