#!/usr/bin/env python
"""small BM25 workload for compute-sanitizer (memcheck / racecheck): every tile size of the filter kernel, column /
run / scanned tokens, a row filter; results checked against the oracle.
   compute-sanitizer --tool racecheck python tools/sanitize_bm25.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from b200rag import _lib  # noqa: E402
from b200rag.bm25 import DeviceBM25, Postings  # noqa: E402
from oracle import numpy_oracle as no  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
docs, n_terms = helpers.zipf_docs(n_docs, 2000, seed=3, lo=10, hi=40)
p = Postings.from_term_ids(docs, n_terms=n_terms)
o = no.CsrBM25(docs)
ix = DeviceBM25(p)
g = np.random.default_rng(1)
df = np.diff(p.term_ptr)
by_df = np.argsort(-df)
qs = [np.concatenate([by_df[g.integers(0, 10, size=3)], by_df[g.integers(10, 200, size=4)], g.integers(0, n_terms, size=3)]).astype(np.int32)
      for _ in range(12)]
allow = g.random(n_docs) < 0.4
bm = np.packbits(allow, bitorder="little")
for tile in (4, 2, 1):
    _lib.set_option("bm25_tile", tile)
    for k in (10, 50):
        for mask, bits in ((None, None), (allow, bm)):
            rows, scores, counts = ix.search_ids(qs, k, bits)
            for i, qt in enumerate(qs[:4]):
                er, es = o.search(qt.tolist(), k, mask)
                assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es), (tile, k, i)
_lib.set_option("bm25_tile", 0)
print("sanitize_bm25 ok", _lib.counters())
