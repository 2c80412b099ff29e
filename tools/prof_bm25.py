#!/usr/bin/env python
"""small BM25 driver for profiling: the C4 corpus of bench.py (1M docs), a few single-query searches and batched ones"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from b200rag import _lib  # noqa: E402
from b200rag.bm25 import DeviceBM25, Postings  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 256
doc_ptr, tokens, n_terms = bench.zipf_tokens(n_docs, 200_000, 1004)
post = Postings.from_flat_tokens(doc_ptr, tokens, n_terms)
ix = DeviceBM25(post)
g = np.random.default_rng(2004)
p = np.arange(1, n_terms + 1, dtype=np.float64) ** (-1.07)
p /= p.sum()
qs = [np.concatenate([g.choice(n_terms, size=g.integers(8, 13), p=p), g.integers(n_terms // 100, n_terms // 10, size=2)]).astype(np.int32)
      for _ in range(Q)]
for q in qs[:6]:
    ix.search_ids([q], 50)
    print("single device ms", float(_lib.last_timings()[0]))
for inl in (0, 1, 0, 1):
    _lib.set_option("bm25_inline_resolve", inl)
    ms = []
    for q in qs[:40]:
        ix.search_ids([q], 50)
        ms.append(float(_lib.last_timings()[0]))
    print("single query, inline resolve", inl, "device ms p50", float(np.median(ms)), "min", min(ms))
_lib.set_option("bm25_inline_resolve", 1)
import os
if os.environ.get("BM25_TILE"):
    _lib.set_option("bm25_tile", int(os.environ["BM25_TILE"]))
for tma in [int(v) for v in os.environ.get("BM25_TMA", "3").split(",")]:
    _lib.set_option("bm25_tma", tma & 1)
    _lib.set_option("bm25_acc16", 1 if tma & 2 else 0)           # 3: 16-bit accumulators, 4 CTAs per SM
    ref = None
    for _ in range(4):
        out = ix.search_ids(qs, 50)
        print("tma", tma, "batch device ms", float(_lib.last_timings()[0]), "per query us", 1e3 * float(_lib.last_timings()[0]) / Q,
              "fallbacks", _lib.counters()["fallbacks"])
    if tma == 0:
        base = out
    elif "base" in globals():
        print("same result as tma 0:", all(np.array_equal(a, b) for a, b in zip(out, base)))
nb, npost = ix.query_bytes(qs)
print("avg filter bytes per query", float(nb.mean()), "avg postings per query", float(npost.mean()),
      "GB/s at last batch", float(nb.sum()) / (float(_lib.last_timings()[0]) * 1e-3) / 1e9)
print("index MB", ix.index_bytes() / 1e6, "bytes/posting", ix.bytes_per_posting())
print("done")
