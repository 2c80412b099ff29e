#!/usr/bin/env python
"""small BM25 driver for profiling: 1M docs, a few single-query and one 64-query search"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    sys.path.insert(0, p)
from b200rag import synth
from b200rag.bm25 import DeviceBM25, Postings

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
docs, n_terms = synth.zipf_corpus(n_docs, 200_000, seed=1004, lo=40, hi=250)
post = Postings.from_term_ids(docs, n_terms=n_terms)
ix = DeviceBM25(post)
g = np.random.default_rng(2004)
p = np.arange(1, n_terms + 1, dtype=np.float64) ** (-1.07)
p /= p.sum()
qs = [np.concatenate([g.choice(n_terms, size=10, p=p), g.integers(n_terms // 100, n_terms // 10, size=2)]).astype(np.int32)
      for _ in range(64)]
for q in qs[:6]:
    ix.search_ids([q], 50)
ix.search_ids(qs, 50)
print("done")
