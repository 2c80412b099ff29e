"""DeviceCollection — drop-in for the `collection` object RAG-DPO injects into
RAGRetriever (src/rag/retriever.py:107-148) and ChunkBM25Index
(src/rag/bm25_index.py:190-239): the chromadb.Collection duck type, cosine
space, with the embedding matrix resident in B200 HBM and exact search.

Methods and result shapes follow the reference's call sites:
  query   src/rag/retriever.py:215-220, 380-385 (+ _parse_chromadb_results :472-494)
  count   src/rag/bm25_index.py:200, app.py:108,116
  get     src/rag/bm25_index.py:211-215, src/processing/ingest_enterprise.py:142,261,290
  add     src/processing/create_chromadb_index.py:374-379, ingest_enterprise.py:241-246
  delete  src/processing/ingest_enterprise.py:272,304
  update  tag_all_chunks.py:215
ids, documents and metadata dicts stay in host Python (row-indexed lists, like
ChunkBM25Index.chunk_ids/chunk_texts/chunk_metadatas, bm25_index.py:184-186).
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from .where import WhereCompiler, match


def l2_normalize_rows(x):
    """Cosine space: rows are normalised at insert and at query time (fp64 norm)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    x64 = x.astype(np.float64)
    n = np.sqrt((x64 ** 2).sum(axis=1, keepdims=True))
    return (x64 / np.maximum(n, 1e-30)).astype(np.float32)


def distance_from_score(score):
    """distance the collection reports: float32(1 - cosine) as a python float"""
    return float(np.float32(1.0 - float(score)))


class DeviceCorpus:
    """Thin owner of a rag_corpus_t handle (the embedding matrix in HBM)."""

    def __init__(self, dim, dtype="bf16", capacity=0):
        self.dim = int(dim)
        self.dtype = _lib.DTYPES[dtype] if isinstance(dtype, str) else int(dtype)
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.rag_corpus_create(C.byref(h), int(capacity), self.dim, self.dtype))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.rag_corpus_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def count(self):
        n = C.c_int64()
        _lib.check(self._L.rag_corpus_count(self._h, C.byref(n)))
        return n.value

    def reserve(self, capacity):
        _lib.check(self._L.rag_corpus_reserve(self._h, int(capacity)))

    def append(self, rows32):
        rows32 = np.ascontiguousarray(rows32, dtype=np.float32)
        assert rows32.ndim == 2 and rows32.shape[1] == self.dim
        _lib.check(self._L.rag_corpus_upload(self._h, self.count(), rows32.shape[0], _lib.ptr(rows32)))

    def overwrite(self, row0, rows32):
        rows32 = np.ascontiguousarray(rows32, dtype=np.float32)
        _lib.check(self._L.rag_corpus_upload(self._h, int(row0), rows32.shape[0], _lib.ptr(rows32)))

    def fill_synthetic(self, seed, nrows, gen_row0=None):
        """append nrows synthetic rows; generator rows gen_row0.. (default: the local row index)"""
        row0 = self.count()
        g0 = row0 if gen_row0 is None else int(gen_row0)
        _lib.check(self._L.rag_corpus_fill_synthetic(self._h, int(seed), g0, row0, int(nrows)))

    def download(self, row0=0, nrows=None):
        nrows = self.count() - row0 if nrows is None else nrows
        out = np.empty((nrows, self.dim), dtype=np.float32)
        _lib.check(self._L.rag_corpus_download(self._h, int(row0), int(nrows), _lib.ptr(out)))
        return out

    def compact(self, keep_rows):
        keep = np.ascontiguousarray(keep_rows, dtype=np.int64)
        _lib.check(self._L.rag_corpus_compact(self._h, _lib.ptr(keep), len(keep)))

    def device_ptr(self):
        p = C.c_void_p()
        _lib.check(self._L.rag_corpus_device_ptr(self._h, C.byref(p)))
        return p.value

    def topk(self, q32, k, allow_bitmap=None, out=None):
        """q32 (B,dim) fp32 host array (already normalised). Returns rows int32
        (B,k) [-1 padded], canonical fp64 scores (B,k), counts int32 (B,)."""
        q32 = np.ascontiguousarray(np.atleast_2d(q32), dtype=np.float32)
        B = q32.shape[0]
        if q32.shape[1] != self.dim:
            raise ValueError(f"query dim {q32.shape[1]} != collection dim {self.dim}")
        if k > _lib.RAG_MAX_K and out is None:
            return self._topk_multipass(q32, int(k), allow_bitmap)
        if out is not None:
            rows, scores, counts = out              # caller-provided (e.g. pinned) result buffers
        else:
            rows = np.empty((B, k), dtype=np.int32)
            scores = np.empty((B, k), dtype=np.float64)
            counts = np.empty(B, dtype=np.int32)
        ab = np.ascontiguousarray(allow_bitmap, dtype=np.uint8) if allow_bitmap is not None else None
        _lib.check(self._L.rag_dense_topk(self._h, _lib.ptr(q32), B, int(k), _lib.ptr(ab), _lib.ptr(rows),
                                          _lib.ptr(scores), _lib.ptr(counts)))
        return rows, scores, counts

    def _topk_multipass(self, q32, k, allow_bitmap):
        """k above the fused select's limit (RAG_MAX_K): exact all the same — each pass takes the next RAG_MAX_K
        rows with the rows already returned masked out (Chroma itself has no such limit: retrieve_candidates with
        n_candidates > 224 must not degrade)."""
        n, B = self.count(), q32.shape[0]
        rows = np.full((B, k), -1, dtype=np.int32)
        scores = np.zeros((B, k), dtype=np.float64)
        counts = np.zeros(B, dtype=np.int32)
        base = (np.unpackbits(np.asarray(allow_bitmap, dtype=np.uint8), bitorder="little")[:n].astype(bool)
                if allow_bitmap is not None else np.ones(n, dtype=bool))
        for b in range(B):
            mask, got = base.copy(), 0
            while got < k:
                kk = min(_lib.RAG_MAX_K, k - got)
                r, s, c = self.topk(q32[b:b + 1], kk, np.packbits(mask, bitorder="little"))
                c0 = int(c[0])
                rows[b, got:got + c0], scores[b, got:got + c0] = r[0, :c0], s[0, :c0]
                mask[r[0, :c0]] = False
                got += c0
                if c0 < kk:
                    break
            counts[b] = got
        return rows, scores, counts

    def topk_dev(self, q_dev_ptr, B, k, out_rows_ptr, out_scores_ptr, out_counts_ptr, allow_dev_ptr=None):
        """device-pointer variant (inputs resident in HBM)."""
        _lib.check(self._L.rag_dense_topk_dev(self._h, q_dev_ptr, int(B), int(k), allow_dev_ptr, out_rows_ptr,
                                              out_scores_ptr, out_counts_ptr))


class DeviceCollection:
    def __init__(self, name="rag_dpo_chunks", dim=1024, dtype="bf16", metadata=None, capacity=0):
        self.name = name
        self.metadata = metadata or {"hnsw:space": "cosine"}
        self.dim = dim
        self.corpus = DeviceCorpus(dim, dtype, capacity)
        self._ids, self._docs, self._metas = [], [], []
        self._pos = {}
        self._where = WhereCompiler()
        self._lock = threading.RLock()     # one cached instance is shared by Streamlit threads (app.py:42)

    # ---- write path -------------------------------------------------------
    def add(self, ids, documents=None, embeddings=None, metadatas=None):
        with self._lock:
            ids = list(ids)
            n = len(ids)
            if embeddings is None:
                raise ValueError("DeviceCollection.add needs embeddings (no embedding function is attached)")
            emb = np.asarray(embeddings, dtype=np.float32)
            if emb.shape != (n, self.dim):
                raise ValueError(f"embeddings shape {emb.shape} != ({n}, {self.dim})")
            documents = list(documents) if documents is not None else [None] * n
            metadatas = list(metadatas) if metadatas is not None else [None] * n
            if len(documents) != n or len(metadatas) != n:
                raise ValueError("ids, documents, embeddings and metadatas must have the same length")
            seen = set()
            for i in ids:
                if i in self._pos or i in seen:
                    raise ValueError(f"duplicate id {i!r}")
                seen.add(i)
            self.corpus.append(l2_normalize_rows(emb))
            base = len(self._ids)
            for j, i in enumerate(ids):
                self._pos[i] = base + j
            self._ids.extend(ids)
            self._docs.extend(documents)
            self._metas.extend(dict(m) if m is not None else None for m in metadatas)
            self._where.invalidate()

    def delete(self, ids=None, where=None):
        with self._lock:
            kill = set()
            if ids is not None:
                kill |= {self._pos[i] for i in ids if i in self._pos}
            if where is not None:
                kill |= {r for r, m in enumerate(self._metas) if match(m, where)}
            if not kill:
                return
            keep = [r for r in range(len(self._ids)) if r not in kill]
            self.corpus.compact(keep)
            self._ids = [self._ids[r] for r in keep]
            self._docs = [self._docs[r] for r in keep]
            self._metas = [self._metas[r] for r in keep]
            self._pos = {i: r for r, i in enumerate(self._ids)}
            self._where.invalidate()

    def update(self, ids, metadatas=None, documents=None, embeddings=None):
        with self._lock:
            for j, i in enumerate(ids):
                r = self._pos[i]
                if metadatas is not None:
                    self._metas[r] = dict(metadatas[j])
                if documents is not None:
                    self._docs[r] = documents[j]
                if embeddings is not None:
                    self.corpus.overwrite(r, l2_normalize_rows(np.asarray(embeddings[j], dtype=np.float32)))
            self._where.invalidate()

    # ---- persistence: flat export / import of the device-resident store -----------
    def save(self, directory):
        """rows.npy (stored values widened to fp32 — exact for bf16/fp16), ids.json, documents.json,
        metadatas.json, manifest.json"""
        import json
        import os
        with self._lock:
            os.makedirs(directory, exist_ok=True)
            np.save(os.path.join(directory, "rows.npy"), self.corpus.download())
            for name, obj in (("ids", self._ids), ("documents", self._docs), ("metadatas", self._metas)):
                with open(os.path.join(directory, f"{name}.json"), "w", encoding="utf-8") as f:
                    json.dump(obj, f, ensure_ascii=False)
            with open(os.path.join(directory, "manifest.json"), "w", encoding="utf-8") as f:
                json.dump({"name": self.name, "dim": self.dim, "dtype": self.corpus.dtype, "count": len(self._ids),
                           "metadata": self.metadata, "format": 1}, f)

    @classmethod
    def load(cls, directory, dtype=None):
        import json
        import os
        with open(os.path.join(directory, "manifest.json"), "r", encoding="utf-8") as f:
            man = json.load(f)
        col = cls(name=man["name"], dim=man["dim"], dtype=man["dtype"] if dtype is None else dtype,
                  metadata=man.get("metadata"), capacity=man["count"])
        rows = np.load(os.path.join(directory, "rows.npy"), mmap_mode="r")
        step = 65536
        for s in range(0, rows.shape[0], step):          # stored values: already normalised and quantised
            col.corpus.append(np.ascontiguousarray(rows[s:s + step], dtype=np.float32))
        for name, attr in (("ids", "_ids"), ("documents", "_docs"), ("metadatas", "_metas")):
            with open(os.path.join(directory, f"{name}.json"), "r", encoding="utf-8") as f:
                setattr(col, attr, json.load(f))
        col._pos = {i: r for r, i in enumerate(col._ids)}
        return col

    # ---- read path --------------------------------------------------------
    def count(self):
        return len(self._ids)

    def get(self, ids=None, where=None, limit=None, offset=None, include=None):
        with self._lock:
            include = include if include is not None else ["documents", "metadatas"]
            if ids is not None:
                rows = sorted(self._pos[i] for i in ids if i in self._pos)
            else:
                rows = range(len(self._ids))
            if where:
                rows = [r for r in rows if match(self._metas[r], where)]
            off = offset or 0
            rows = list(rows[off:off + limit] if limit is not None else rows[off:])
            out = {"ids": [self._ids[r] for r in rows],
                   "documents": [self._docs[r] for r in rows] if "documents" in include else None,
                   "metadatas": [self._metas[r] for r in rows] if "metadatas" in include else None,
                   "embeddings": None}
            if "embeddings" in include:
                if rows and rows == list(range(rows[0], rows[-1] + 1)):
                    out["embeddings"] = self.corpus.download(rows[0], len(rows))
                else:
                    out["embeddings"] = (np.stack([self.corpus.download(r, 1)[0] for r in rows])
                                         if rows else np.zeros((0, self.dim), np.float32))
            return out

    def doc_filter_bitmap(self, doc_paths):
        """row bitmap of the chunks whose metadata document_path is in doc_paths"""
        key = frozenset(doc_paths)
        mask = np.fromiter(((m or {}).get("document_path", "") in key for m in self._metas), dtype=bool,
                           count=len(self._metas))
        return np.packbits(mask, bitorder="little"), int(mask.sum())

    def query_rows(self, query_embeddings, n_results=10, where=None, doc_filter=None):
        """Batched array-level query: rows int32 (B,k), fp64 cosine (B,k), counts (B,).
        doc_filter (opt-in, NOT what the reference does): restrict the search to the chunks of these documents
        BEFORE the top-k, instead of post-filtering the results (src/rag/retriever.py:393-398)."""
        with self._lock:
            q = l2_normalize_rows(np.asarray(query_embeddings, dtype=np.float32))
            n = len(self._ids)
            bitmap, allowed = self._where.compile(self._metas, where)
            if doc_filter is not None:
                fb, fa = self.doc_filter_bitmap(doc_filter)
                bitmap = fb if bitmap is None else np.bitwise_and(bitmap, fb)
                allowed = int(np.unpackbits(bitmap, bitorder="little")[:n].sum())
            k = min(int(n_results), n)
            if k <= 0 or allowed == 0:
                B = q.shape[0]
                return (np.full((B, 0), -1, np.int32), np.zeros((B, 0), np.float64), np.zeros(B, np.int32))
            return self.corpus.topk(q, k, bitmap)

    def query(self, query_embeddings, n_results=10, where=None, include=None):
        include = include if include is not None else ["documents", "metadatas", "distances"]
        with self._lock:
            rows, scores, counts = self.query_rows(query_embeddings, n_results, where)
            out = {"ids": [], "documents": [] if "documents" in include else None,
                   "metadatas": [] if "metadatas" in include else None,
                   "distances": [] if "distances" in include else None}
            for b in range(rows.shape[0]):
                rr = rows[b, :counts[b]].tolist()
                out["ids"].append([self._ids[r] for r in rr])
                if out["documents"] is not None:
                    out["documents"].append([self._docs[r] for r in rr])
                if out["metadatas"] is not None:
                    out["metadatas"].append([self._metas[r] for r in rr])
                if out["distances"] is not None:
                    out["distances"].append([distance_from_score(s) for s in scores[b, :counts[b]]])
            return out
