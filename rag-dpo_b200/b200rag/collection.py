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
On the device next to the rows: a tombstone bitmap (delete is O(rows deleted),
row numbers never move until a compaction) and the metadata as dictionary-coded
int32 columns, so that a `where` filter is evaluated by a device kernel into a
cached row bitmap instead of a Python loop over N dicts per query.

n_shards > 1: ONE process drives several GPUs (the reference is one Streamlit
process holding one cached collection, app.py:42-119): rows are spread
block-cyclically over the GPUs, every query runs on all of them and the
per-GPU top-k lists are merged over NVLink peer memory inside the library.
"""
import ctypes as C
import logging
import threading

import numpy as np

from . import _lib
from .where import ColumnCodes, Unsupported, compile_where, match

logger = logging.getLogger(__name__)


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
    """Thin owner of a rag_corpus_t handle (the embedding matrix in HBM, on one GPU or sharded over several)."""

    def __init__(self, dim, dtype="bf16", capacity=0, n_shards=1, devices=None):
        self.dim = int(dim)
        self.dtype = _lib.DTYPES[dtype] if isinstance(dtype, str) else int(dtype)
        self.n_shards = int(n_shards)
        self._L = _lib.lib()
        h = C.c_void_p()
        if self.n_shards > 1:
            _lib.ensure_slots(self.n_shards, devices)
            _lib.check(self._L.rag_corpus_create_sharded(C.byref(h), int(capacity), self.dim, self.dtype, self.n_shards))
        else:
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
        """rows including tombstones"""
        n = C.c_int64()
        _lib.check(self._L.rag_corpus_count(self._h, C.byref(n)))
        return n.value

    def live_count(self):
        n = C.c_int64()
        _lib.check(self._L.rag_corpus_live_count(self._h, C.byref(n)))
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
        """append nrows synthetic rows; generator rows gen_row0.. (default: the row index)"""
        row0 = self.count()
        g0 = row0 if gen_row0 is None else int(gen_row0)
        _lib.check(self._L.rag_corpus_fill_synthetic(self._h, int(seed), g0, row0, int(nrows)))

    def download(self, row0=0, nrows=None):
        nrows = self.count() - row0 if nrows is None else nrows
        out = np.empty((nrows, self.dim), dtype=np.float32)
        _lib.check(self._L.rag_corpus_download(self._h, int(row0), int(nrows), _lib.ptr(out)))
        return out

    def delete_rows(self, rows):
        """tombstones: O(len(rows)); the rows must be alive"""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        _lib.check(self._L.rag_corpus_delete_rows(self._h, _lib.ptr(rows), len(rows)))

    def set_codes(self, column, row0, codes):
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        _lib.check(self._L.rag_corpus_set_codes(self._h, int(column), int(row0), len(codes), _lib.ptr(codes)))

    def compact(self, keep_rows):
        keep = np.ascontiguousarray(keep_rows, dtype=np.int64)
        _lib.check(self._L.rag_corpus_compact(self._h, _lib.ptr(keep), len(keep)))

    def device_ptr(self):
        p = C.c_void_p()
        _lib.check(self._L.rag_corpus_device_ptr(self._h, C.byref(p)))
        return p.value

    def _prep(self, q32, k, out):
        q32 = np.ascontiguousarray(np.atleast_2d(q32), dtype=np.float32)
        B = q32.shape[0]
        if q32.shape[1] != self.dim:
            raise ValueError(f"query dim {q32.shape[1]} != collection dim {self.dim}")
        if out is not None:
            rows, scores, counts = out              # caller-provided (e.g. pinned) result buffers
        else:
            rows = np.empty((B, k), dtype=np.int32)
            scores = np.empty((B, k), dtype=np.float64)
            counts = np.empty(B, dtype=np.int32)
        return q32, B, rows, scores, counts

    def topk(self, q32, k, allow_bitmap=None, out=None):
        """q32 (B,dim) fp32 host array (already normalised). Returns rows int32
        (B,k) [-1 padded], canonical fp64 scores (B,k), counts int32 (B,)."""
        if k > _lib.RAG_MAX_K and out is None:
            q32 = np.ascontiguousarray(np.atleast_2d(q32), dtype=np.float32)
            if q32.shape[1] != self.dim:
                raise ValueError(f"query dim {q32.shape[1]} != collection dim {self.dim}")
            return self._topk_multipass(q32, int(k), allow_bitmap)
        q32, B, rows, scores, counts = self._prep(q32, k, out)
        ab = np.ascontiguousarray(allow_bitmap, dtype=np.uint8) if allow_bitmap is not None else None
        _lib.check(self._L.rag_dense_topk(self._h, _lib.ptr(q32), B, int(k), _lib.ptr(ab), _lib.ptr(rows),
                                          _lib.ptr(scores), _lib.ptr(counts)))
        return rows, scores, counts

    def topk_where(self, q32, k, prog, out=None):
        """same, filtered by a compiled `where` program (where.compile_where): evaluated on the device, cached"""
        if prog is None or len(prog) == 0:
            return self.topk(q32, k, out=out)
        if k > _lib.RAG_MAX_K:
            raise Unsupported("k above RAG_MAX_K takes the bitmap path")
        q32, B, rows, scores, counts = self._prep(q32, k, out)
        prog = np.ascontiguousarray(prog, dtype=np.int32)
        _lib.check(self._L.rag_dense_topk_where(self._h, _lib.ptr(q32), B, int(k), _lib.ptr(prog), len(prog),
                                                _lib.ptr(rows), _lib.ptr(scores), _lib.ptr(counts)))
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

    def debug_last_candidates(self, n_queries, cap):
        """test hook: the tensor-core filter's candidate lists of the last call -> (rows (B,cap) int64 [-1 padded],
        fp32 filter scores (B,cap), counts (B,), eps_rel)"""
        keys = np.zeros((n_queries, cap), dtype=np.uint64)
        counts = np.zeros(n_queries, dtype=np.int32)
        eps = C.c_double(0.0)
        _lib.check(self._L.rag_debug_last_candidates(self._h, int(n_queries), _lib.ptr(keys), int(cap), _lib.ptr(counts),
                                                     C.byref(eps)))
        o = (keys >> np.uint64(32)).astype(np.uint32)
        u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o)
        scores = u.astype(np.uint32).view(np.float32)
        rows = (~keys.astype(np.uint32)).astype(np.int64)          # low word = ~row
        live = np.arange(cap)[None, :] < counts[:, None]
        rows[~live] = -1
        return rows, scores, counts, float(eps.value)

    def topk_dev(self, q_dev_ptr, B, k, out_rows_ptr, out_scores_ptr, out_counts_ptr, allow_dev_ptr=None):
        """device-pointer variant (inputs resident in HBM): STREAM-ORDERED, only queues work."""
        _lib.check(self._L.rag_dense_topk_dev(self._h, q_dev_ptr, int(B), int(k), allow_dev_ptr, out_rows_ptr,
                                              out_scores_ptr, out_counts_ptr))


class DeviceCollection:
    COMPACT_MIN_DEAD = 4096          # automatic compaction: at least this many tombstones ...
    COMPACT_FRACTION = 0.25          # ... and this fraction of the rows

    def __init__(self, name="rag_dpo_chunks", dim=1024, dtype="bf16", metadata=None, capacity=0, n_shards=1,
                 devices=None):
        self.name = name
        self.metadata = metadata or {"hnsw:space": "cosine"}
        self.dim = dim
        self.corpus = DeviceCorpus(dim, dtype, capacity, n_shards=n_shards, devices=devices)
        self._ids, self._docs, self._metas = [], [], []       # row-indexed; None at tombstones
        self._pos = {}
        self._n_dead = 0
        self._cols = ColumnCodes()
        self._live_cache = None          # ascending live rows (lazily rebuilt)
        self._prog_cache = {}
        self.mutation_version = 0
        self._lock = threading.RLock()     # one cached instance is shared by Streamlit threads (app.py:42)

    def _mutated(self):
        self.mutation_version += 1
        self._live_cache = None
        self._prog_cache.clear()

    def _send_codes(self, row0, metadatas):
        for col, codes in self._cols.encode_batch(metadatas).items():
            self.corpus.set_codes(col, row0, codes)

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
            self._send_codes(base, self._metas[base:])
            self._mutated()

    def delete(self, ids=None, where=None):
        """tombstones on the device (O(rows deleted)); row numbers stay put until a compaction"""
        with self._lock:
            kill = set()
            if ids is not None:
                kill |= {self._pos[i] for i in ids if i in self._pos}
            if where is not None:
                kill |= {r for r in self._live_rows().tolist() if match(self._metas[r], where)}
            if not kill:
                return
            rows = sorted(kill)
            self.corpus.delete_rows(rows)
            for r in rows:
                del self._pos[self._ids[r]]
                self._ids[r] = self._docs[r] = self._metas[r] = None
            self._n_dead += len(rows)
            self._mutated()
            if (self.corpus.n_shards == 1 and self._n_dead >= self.COMPACT_MIN_DEAD
                    and self._n_dead >= self.COMPACT_FRACTION * len(self._ids)):
                self.compact()

    def compact(self):
        """physically drop the tombstoned rows (single-GPU collections): rows are renumbered in order"""
        with self._lock:
            if self._n_dead == 0:
                return
            keep = self._live_rows()
            self.corpus.compact(keep)
            self._ids = [self._ids[r] for r in keep.tolist()]
            self._docs = [self._docs[r] for r in keep.tolist()]
            self._metas = [self._metas[r] for r in keep.tolist()]
            self._pos = {i: r for r, i in enumerate(self._ids)}
            self._n_dead = 0
            self._send_codes(0, self._metas)
            self._mutated()

    def update(self, ids, metadatas=None, documents=None, embeddings=None):
        with self._lock:
            for j, i in enumerate(ids):
                r = self._pos[i]
                if metadatas is not None:
                    self._metas[r] = dict(metadatas[j])
                    codes = self._cols.encode_batch([self._metas[r]])
                    for col in range(len(self._cols.codes)):       # keys the new metadata lacks become "missing"
                        self.corpus.set_codes(col, r, codes.get(col, np.full(1, -1, np.int32)))
                if documents is not None:
                    self._docs[r] = documents[j]
                if embeddings is not None:
                    self.corpus.overwrite(r, l2_normalize_rows(np.asarray(embeddings[j], dtype=np.float32)))
            self._mutated()

    # ---- persistence: flat export / import of the device-resident store -----------
    def save(self, directory):
        """rows.npy (stored values widened to fp32 — exact for bf16/fp16), ids.json, documents.json,
        metadatas.json, manifest.json; tombstoned rows are not written"""
        import json
        import os
        with self._lock:
            os.makedirs(directory, exist_ok=True)
            live = self._live_rows()
            rows = self.corpus.download()
            np.save(os.path.join(directory, "rows.npy"), rows if self._n_dead == 0 else rows[live])
            for name, obj in (("ids", self._ids), ("documents", self._docs), ("metadatas", self._metas)):
                with open(os.path.join(directory, f"{name}.json"), "w", encoding="utf-8") as f:
                    json.dump([obj[r] for r in live.tolist()], f, ensure_ascii=False)
            with open(os.path.join(directory, "manifest.json"), "w", encoding="utf-8") as f:
                json.dump({"name": self.name, "dim": self.dim, "dtype": self.corpus.dtype, "count": int(len(live)),
                           "metadata": self.metadata, "format": 1}, f)

    @classmethod
    def load(cls, directory, dtype=None, n_shards=1, devices=None):
        import json
        import os
        with open(os.path.join(directory, "manifest.json"), "r", encoding="utf-8") as f:
            man = json.load(f)
        col = cls(name=man["name"], dim=man["dim"], dtype=man["dtype"] if dtype is None else dtype,
                  metadata=man.get("metadata"), capacity=man["count"], n_shards=n_shards, devices=devices)
        rows = np.load(os.path.join(directory, "rows.npy"), mmap_mode="r")
        step = 65536
        for s in range(0, rows.shape[0], step):          # stored values: already normalised and quantised
            col.corpus.append(np.ascontiguousarray(rows[s:s + step], dtype=np.float32))
        for name, attr in (("ids", "_ids"), ("documents", "_docs"), ("metadatas", "_metas")):
            with open(os.path.join(directory, f"{name}.json"), "r", encoding="utf-8") as f:
                setattr(col, attr, json.load(f))
        col._pos = {i: r for r, i in enumerate(col._ids)}
        for s in range(0, len(col._metas), step):
            col._send_codes(s, col._metas[s:s + step])
        col._mutated()
        return col

    # ---- read path --------------------------------------------------------
    def count(self):
        return len(self._ids) - self._n_dead

    def _live_rows(self):
        if self._live_cache is None:
            if self._n_dead == 0:
                self._live_cache = np.arange(len(self._ids), dtype=np.int64)
            else:
                self._live_cache = np.fromiter((r for r, i in enumerate(self._ids) if i is not None), dtype=np.int64)
        return self._live_cache

    def get(self, ids=None, where=None, limit=None, offset=None, include=None):
        with self._lock:
            include = include if include is not None else ["documents", "metadatas"]
            if ids is not None:
                rows = sorted(self._pos[i] for i in ids if i in self._pos)
            else:
                rows = self._live_rows().tolist() if (where or self._n_dead) else range(len(self._ids))
            if where:
                rows = [r for r in rows if match(self._metas[r], where)]
            off = offset or 0
            rows = list(rows[off:off + limit] if limit is not None else rows[off:])
            out = {"ids": [self._ids[r] for r in rows],
                   "documents": [self._docs[r] for r in rows] if "documents" in include else None,
                   "metadatas": [dict(self._metas[r]) if self._metas[r] is not None else None for r in rows]
                   if "metadatas" in include else None,
                   "embeddings": None}
            if "embeddings" in include:
                if rows and rows == list(range(rows[0], rows[-1] + 1)):
                    out["embeddings"] = self.corpus.download(rows[0], len(rows))
                else:
                    out["embeddings"] = (np.stack([self.corpus.download(r, 1)[0] for r in rows])
                                         if rows else np.zeros((0, self.dim), np.float32))
            return out

    def doc_filter_bitmap(self, doc_paths):
        """row bitmap of the live chunks whose metadata document_path is in doc_paths (host form of the N4 filter)"""
        key = frozenset(doc_paths)
        mask = np.fromiter(((m or {}).get("document_path", "") in key for m in self._metas), dtype=bool,
                           count=len(self._metas))
        return np.packbits(mask, bitorder="little"), int(mask.sum())

    def _program(self, where, doc_filter):
        """compiled predicate for (where, doc_filter), cached until the next mutation; None = no filter;
        raises Unsupported when only the host evaluator covers it"""
        import json
        key = (json.dumps(where, sort_keys=True, default=str) if where else "",
               frozenset(doc_filter) if doc_filter is not None else None)
        hit = self._prog_cache.get(key)
        if hit is None:
            try:
                hit = ("ok", compile_where(where, self._cols, doc_filter))
            except Unsupported as e:
                hit = ("host", str(e))
            if len(self._prog_cache) > 64:
                self._prog_cache.clear()
            self._prog_cache[key] = hit
        if hit[0] == "host":
            raise Unsupported(hit[1])
        return hit[1]

    def query_rows(self, query_embeddings, n_results=10, where=None, doc_filter=None):
        """Batched array-level query: rows int32 (B,k), fp64 cosine (B,k), counts (B,).
        doc_filter (opt-in, NOT what the reference does): restrict the search to the chunks of these documents
        BEFORE the top-k, instead of post-filtering the results (src/rag/retriever.py:393-398)."""
        with self._lock:
            q = l2_normalize_rows(np.asarray(query_embeddings, dtype=np.float32))
            B = q.shape[0]
            k = min(int(n_results), self.count())
            if k <= 0:
                return (np.full((B, 0), -1, np.int32), np.zeros((B, 0), np.float64), np.zeros(B, np.int32))
            if k <= _lib.RAG_MAX_K:
                try:
                    prog = self._program(where, doc_filter)
                    return self.corpus.topk_where(q, k, prog)
                except Unsupported:
                    pass
            # host evaluator: operators outside the device program ($gt, ...), or k above the fused select's limit
            if not where and doc_filter is None:
                return self.corpus.topk(q, k)
            paths = frozenset(doc_filter) if doc_filter is not None else None
            mask = np.fromiter((m is not None and match(m, where) and
                                (paths is None or m.get("document_path", "") in paths) for m in self._metas),
                               dtype=bool, count=len(self._metas))
            if not mask.any():
                return (np.full((B, 0), -1, np.int32), np.zeros((B, 0), np.float64), np.zeros(B, np.int32))
            return self.corpus.topk(q, k, np.packbits(mask, bitorder="little"))

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
                    # fresh dicts: the caller mutates what it gets back (retriever.py:252-256)
                    out["metadatas"].append([dict(self._metas[r]) if self._metas[r] is not None else None for r in rr])
                if out["distances"] is not None:
                    out["distances"].append([distance_from_score(s) for s in scores[b, :counts[b]]])
            return out
