"""ORACLE — test infrastructure, not product code.

numpy restatement of the arithmetic on the retrieval hot path of
MatJoss/RAG-DPO (citations relative to /root/reference):

  dense      collection.query(...) as called at src/rag/retriever.py:215-220,
             380-385 against a collection created with "hnsw:space": "cosine"
             (src/processing/create_chromadb_index.py:100-106): distance =
             1 - <q^, x^>, ascending, rows failing ``where`` excluded BEFORE
             selection.  chromadb==1.4.1 itself is absent (PARITY UNPINNED for
             the third-party HNSW; exact search is a recall-1.0 superset).
  bm25       rank-bm25==0.2.2 BM25Okapi.get_scores (oracle/rank_bm25.py) plus
             the reference-owned select in src/rag/bm25_index.py:265-279.
  rrf        reciprocal_rank_fusion, src/rag/retriever.py:66-90, and the fusion
             tail :454-467 (stable sort, first-seen order).

The canonical dense score is the fp64 sum of the exact
products of the stored values in a FIXED order, so GPU and oracle agree
bit-for-bit:  32 partial sums; p[l] owns the 8-element groups g = l, l+32,
l+64, ... (elements 8g..8g+7) and adds their products in ascending element
order; then p[l] += p[l+off] for off = 16,8,4,2,1.  (A group is one 16-byte
vector of a bf16/fp16 row, two of an fp32 row: the order a GPU warp reads in.)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import anything under oracle/.
"""
import json
import math

import numpy as np

DT_F32, DT_BF16, DT_F16 = 0, 1, 2


# --------------------------------------------------------------------------
# storage dtypes
# --------------------------------------------------------------------------
def f32_to_bf16_bits(x):
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def quantize(x32, dtype):
    """fp32 array -> the values actually stored for `dtype`, as fp32."""
    x32 = np.ascontiguousarray(x32, dtype=np.float32)
    if dtype == DT_F32:
        return x32
    if dtype == DT_BF16:
        return bf16_bits_to_f32(f32_to_bf16_bits(x32))
    if dtype == DT_F16:
        return x32.astype(np.float16).astype(np.float32)
    raise ValueError(dtype)


def l2_normalize_rows(x):
    """Row-normalise like a cosine-space index does at insert/query time:
    fp64 norm, fp32 result."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    n = np.maximum(n, 1e-30)
    return (x.astype(np.float64) / n).astype(np.float32)


# --------------------------------------------------------------------------
# dense: canonical fp64 score + exact top-k
# --------------------------------------------------------------------------
def canonical_scores(q32, x_stored32):
    """q32: (D,) fp32;  x_stored32: (n, D) fp32 holding the stored values.
    Returns (n,) fp64 canonical scores."""
    q = np.asarray(q32, dtype=np.float32).astype(np.float64)
    x = np.atleast_2d(x_stored32)
    n, d = x.shape
    assert d % 8 == 0
    groups = d // 8
    p = np.zeros((n, 32), dtype=np.float64)
    for c in range((groups + 31) // 32):            # lane l takes group 32c + l of this chunk
        w = min(32, groups - 32 * c)
        for i in range(8):
            cols = slice(256 * c + i, 256 * c + 8 * w, 8)
            p[:, :w] += x[:, cols].astype(np.float64) * q[cols]
    off = 16
    while off >= 1:
        p[:, :off] += p[:, off:2 * off]
        off //= 2
    return p[:, 0].copy()


def dense_topk(q32, x_stored32, k, allow=None, chunk=65536):
    """Exact top-k by (canonical score desc, row asc).  Returns (rows int64,
    scores fp64)."""
    n = x_stored32.shape[0]
    scores = np.empty(n, dtype=np.float64)
    for s in range(0, n, chunk):
        scores[s:s + chunk] = canonical_scores(q32, x_stored32[s:s + chunk])
    rows = np.arange(n, dtype=np.int64)
    if allow is not None:
        rows = rows[np.asarray(allow, dtype=bool)]
    sc = scores[rows]
    order = np.lexsort((rows, -sc))[:k]
    return rows[order], sc[order]


def distance_from_score(score64):
    """What the collection reports: float32(1 - cos) widened to a python float."""
    return float(np.float32(1.0 - float(score64)))


# --------------------------------------------------------------------------
# where-filter evaluator (Chroma grammar actually emitted by the reference:
# src/rag/pipeline.py:59-69, pages/1_*Chat.py:247, test_rag.py:145,
# src/processing/ingest_enterprise.py:291-294)
# --------------------------------------------------------------------------
def where_match(meta, where):
    if not where:
        return True
    for key, cond in where.items():
        if key == "$and":
            if not all(where_match(meta, w) for w in cond):
                return False
        elif key == "$or":
            if not any(where_match(meta, w) for w in cond):
                return False
        elif isinstance(cond, dict):
            for op, val in cond.items():
                present = key in meta
                v = meta.get(key)
                if op == "$eq":
                    ok = present and type(v) is type(val) and v == val
                elif op == "$ne":
                    ok = (not present) or not (type(v) is type(val) and v == val)
                elif op == "$in":
                    ok = present and any(type(v) is type(x) and v == x for x in val)
                elif op == "$nin":
                    ok = (not present) or not any(type(v) is type(x) and v == x for x in val)
                elif op in ("$gt", "$gte", "$lt", "$lte"):
                    if not present or isinstance(v, (str, bool)) or isinstance(val, (str, bool)):
                        ok = False
                    else:
                        ok = {"$gt": v > val, "$gte": v >= val, "$lt": v < val, "$lte": v <= val}[op]
                else:
                    raise ValueError(f"unsupported where operator {op}")
                if not ok:
                    return False
        else:
            if not (key in meta and type(meta[key]) is type(cond) and meta[key] == cond):
                return False
    return True


class ExactCollection:
    """Duck-typed stand-in for chromadb.Collection (cosine space) doing the
    canonical exact search.  Methods = the ones the reference calls
    (SURVEY.md §8b)."""

    def __init__(self, name="rag_dpo_chunks", dim=1024, dtype=DT_F32, metadata=None):
        self.name = name
        self.metadata = metadata or {"hnsw:space": "cosine"}
        self.dim = dim
        self.dtype = dtype
        self._ids = []
        self._docs = []
        self._metas = []
        self._x = np.zeros((0, dim), dtype=np.float32)   # stored values, as fp32
        self._pos = {}

    # -- write path ---------------------------------------------------------
    def add(self, ids, documents=None, embeddings=None, metadatas=None):
        n = len(ids)
        documents = documents if documents is not None else [None] * n
        metadatas = metadatas if metadatas is not None else [None] * n
        emb = quantize(l2_normalize_rows(np.asarray(embeddings, dtype=np.float32)), self.dtype)
        for i in ids:
            if i in self._pos:
                raise ValueError(f"duplicate id {i}")
        base = len(self._ids)
        for j, i in enumerate(ids):
            self._pos[i] = base + j
        self._ids.extend(ids)
        self._docs.extend(documents)
        self._metas.extend(dict(m) if m is not None else None for m in metadatas)
        self._x = np.concatenate([self._x, emb], axis=0)

    def delete(self, ids=None, where=None):
        kill = set()
        if ids is not None:
            kill |= {self._pos[i] for i in ids if i in self._pos}
        if where is not None:
            kill |= {r for r, m in enumerate(self._metas) if where_match(m or {}, where)}
        keep = [r for r in range(len(self._ids)) if r not in kill]
        self._ids = [self._ids[r] for r in keep]
        self._docs = [self._docs[r] for r in keep]
        self._metas = [self._metas[r] for r in keep]
        self._x = self._x[keep]
        self._pos = {i: r for r, i in enumerate(self._ids)}

    def update(self, ids, metadatas=None, documents=None, embeddings=None):
        for j, i in enumerate(ids):
            r = self._pos[i]
            if metadatas is not None:
                self._metas[r] = dict(metadatas[j])
            if documents is not None:
                self._docs[r] = documents[j]
            if embeddings is not None:
                self._x[r] = quantize(l2_normalize_rows(np.asarray(embeddings[j], dtype=np.float32)), self.dtype)[0]

    # -- read path ----------------------------------------------------------
    def count(self):
        return len(self._ids)

    def get(self, ids=None, where=None, limit=None, offset=None, include=None):
        include = include if include is not None else ["documents", "metadatas"]
        if ids is not None:
            rows = [self._pos[i] for i in ids if i in self._pos]
            rows.sort()
        else:
            rows = list(range(len(self._ids)))
        if where:
            rows = [r for r in rows if where_match(self._metas[r] or {}, where)]
        off = offset or 0
        rows = rows[off:off + limit] if limit is not None else rows[off:]
        out = {"ids": [self._ids[r] for r in rows],
               "documents": [self._docs[r] for r in rows] if "documents" in include else None,
               "metadatas": [self._metas[r] for r in rows] if "metadatas" in include else None,
               "embeddings": self._x[rows].copy() if "embeddings" in include else None}
        return out

    def query(self, query_embeddings, n_results=10, where=None, include=None):
        include = include if include is not None else ["documents", "metadatas", "distances"]
        allow = None
        if where:
            allow = np.fromiter((where_match(m or {}, where) for m in self._metas), dtype=bool,
                                count=len(self._metas))
        out = {"ids": [], "documents": [] if "documents" in include else None,
               "metadatas": [] if "metadatas" in include else None,
               "distances": [] if "distances" in include else None}
        qn = l2_normalize_rows(np.asarray(query_embeddings, dtype=np.float32))
        for q in qn:
            rows, sc = dense_topk(q, self._x, n_results, allow)
            out["ids"].append([self._ids[r] for r in rows])
            if out["documents"] is not None:
                out["documents"].append([self._docs[r] for r in rows])
            if out["metadatas"] is not None:
                out["metadatas"].append([self._metas[r] for r in rows])
            if out["distances"] is not None:
                out["distances"].append([distance_from_score(s) for s in sc])
        return out


# --------------------------------------------------------------------------
# BM25 over CSR postings (fast restatement of oracle/rank_bm25.py)
# --------------------------------------------------------------------------
class CsrBM25:
    """Same arithmetic as rank_bm25.BM25Okapi (oracle/rank_bm25.py) but over
    integer term ids and CSR postings so that 1M-document cases finish in
    seconds.  tests/test_oracle.py checks it bit-for-bit against the dict
    version on small corpora."""

    def __init__(self, docs_term_ids, k1=1.5, b=0.75, epsilon=0.25):
        """docs_term_ids: list of 1-D int arrays (token ids in document order).
        Vocabulary ids must be assigned in first-seen order for the idf sum
        order to match the dict version."""
        self.k1, self.b, self.epsilon = k1, b, epsilon
        n = len(docs_term_ids)
        self.n = n
        self.doc_len = np.array([len(d) for d in docs_term_ids], dtype=np.int64)
        self.avgdl = int(self.doc_len.sum()) / n
        rows, terms, tfs = [], [], []
        for r, d in enumerate(docs_term_ids):
            if len(d) == 0:
                continue
            t, c = np.unique(np.asarray(d, dtype=np.int64), return_counts=True)
            rows.append(np.full(len(t), r, dtype=np.int64))
            terms.append(t)
            tfs.append(c)
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        terms = np.concatenate(terms) if terms else np.zeros(0, np.int64)
        tfs = np.concatenate(tfs) if tfs else np.zeros(0, np.int64)
        v = int(terms.max()) + 1 if len(terms) else 0
        order = np.lexsort((rows, terms))
        self.post_row = rows[order].astype(np.int32)
        self.post_tf = tfs[order].astype(np.int32)
        df = np.bincount(terms, minlength=v)
        self.term_ptr = np.zeros(v + 1, dtype=np.int64)
        np.cumsum(df, out=self.term_ptr[1:])
        self.df = df
        self.idf = self._calc_idf(df)

    def _calc_idf(self, df):
        idf = np.zeros(len(df), dtype=np.float64)
        idf_sum = 0
        present = 0
        for t in range(len(df)):           # vocabulary (first-seen) order
            if df[t] == 0:
                continue
            v = math.log(self.n - int(df[t]) + 0.5) - math.log(int(df[t]) + 0.5)
            idf[t] = v
            idf_sum += v
            present += 1
        self.average_idf = idf_sum / present
        floor = self.epsilon * self.average_idf
        idf[(idf < 0) & (df > 0)] = floor
        return idf

    def get_scores(self, query_term_ids):
        score = np.zeros(self.n, dtype=np.float64)
        dl = self.doc_len
        for t in query_term_ids:
            if t < 0 or t >= len(self.df) or self.df[t] == 0:
                continue                   # unknown token: (idf.get(q) or 0) * 0 adds +0.0
            w = self.idf[t]
            if w == 0:
                continue
            lo, hi = self.term_ptr[t], self.term_ptr[t + 1]
            r = self.post_row[lo:hi]
            tf = self.post_tf[lo:hi].astype(np.int64)
            score[r] += w * (tf * (self.k1 + 1) / (tf + self.k1 * (1 - self.b + self.b * dl[r] / self.avgdl)))
        return score

    def search(self, query_term_ids, top_k, allow=None):
        """Reference select (src/rag/bm25_index.py:267-279): drop <= 0, apply the
        row filter, stable sort descending, first top_k."""
        s = self.get_scores(query_term_ids)
        rows = np.nonzero(s > 0)[0]
        if allow is not None:
            rows = rows[np.asarray(allow, dtype=bool)[rows]]
        order = np.lexsort((rows, -s[rows]))[:top_k]
        return rows[order].astype(np.int64), s[rows][order]


# --------------------------------------------------------------------------
# RRF (src/rag/retriever.py:66-90) over integer ids + the fusion tail :454-467
# --------------------------------------------------------------------------
def rrf_fuse(rankings, weights=None, k=60, top=None):
    """rankings: list of lists of hashable ids.  Returns (ids, scores) sorted by
    (score desc, first-seen order) — what `sorted(chunk_map.values(),
    key=hybrid_score, reverse=True)[:top]` yields."""
    if weights is None:
        weights = [1.0] * len(rankings)
    scores = {}
    for ranking, w in zip(rankings, weights):
        for rank, i in enumerate(ranking):
            scores[i] = scores.get(i, 0.0) + w / (k + rank + 1)
    ids = list(scores.keys())                       # dict order == first-seen order
    ids.sort(key=lambda i: scores[i], reverse=True)  # stable
    if top is not None:
        ids = ids[:top]
    return ids, [scores[i] for i in ids]


# --------------------------------------------------------------------------
# rerank select: the tail of CrossEncoderReranker.rerank (src/rag/reranker.py:172-211)
# --------------------------------------------------------------------------
def rerank_select(scores, boosts, top_k, min_score):
    """scores: model scores (fp32) in candidate order, boosts: per-candidate topic boost (0.0 = none).
    Returns (positions in the candidate list, final scores) of what the reference returns."""
    final = []
    for s, b in zip(scores, boosts):
        f = float(np.float32(s))
        if b > 0:
            f += float(b)
        final.append(f)
    order = list(range(len(final)))
    order.sort(key=lambda i: final[i], reverse=True)            # stable
    result = [i for i in order[:top_k] if final[i] >= min_score]
    if len(result) < 3 and len(order) >= 3:
        result = order[:3]
    return result, [final[i] for i in result]


def load_json(path):
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)
