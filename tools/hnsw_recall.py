#!/usr/bin/env python
"""recall@k of the reference's approximate vector store against the exact search, CPU only (no GPU needed: the
GPU path is bit-identical to the exact oracle, see tests/).  chromadb 1.4.1 / hnswlib are not installable here, so
the store is the restated HNSW of oracle/hnsw.c with chromadb's defaults (M=16, ef_construction=100, ef_search=100,
cosine space) — PARITY UNPINNED, a statistical twin.  Corpora: BASELINE config 1 (50k x 1024 synthetic i.i.d. unit
rows, the 48 x 4 query variants of tools/bench_c1.py) and a clustered twin (10 chunks per document around a
document vector), which is what real chunk embeddings look like.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def recall_case(name, x, q, ks, ef_search):
    from oracle import c_oracle
    t0 = time.perf_counter()
    ix = c_oracle.HnswIndex(x, M=16, ef_construction=100)
    build_s = time.perf_counter() - t0
    s = q @ x.T                                         # exact fp32 scores; top-k sets only (ties are measure-zero)
    out = {"corpus": name, "rows": int(x.shape[0]), "queries": int(q.shape[0]), "hnsw_build_s": round(build_s, 1)}
    for k in ks:
        exact = np.argsort(-s, axis=1, kind="stable")[:, :k]
        t0 = time.perf_counter()
        ids, _ = ix.query(q, k, ef_search)
        ms = 1e3 * (time.perf_counter() - t0) / len(q)
        rec = float(np.mean([len(set(ids[i].tolist()) & set(exact[i].tolist())) / k for i in range(len(q))]))
        out[f"recall@{k}"] = round(rec, 4)
        out[f"hnsw_query_ms@{k}"] = round(ms, 3)
    ix.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=50_000)
    ap.add_argument("--ef-search", type=int, default=100)
    args = ap.parse_args()
    from b200rag import synth
    from oracle import numpy_oracle as no
    n, d = args.chunks, 1024
    x = no.l2_normalize_rows(synth.synth_rows(1001, 0, n, d))
    q = synth.unit_queries(48 * 4, d, 2001)
    res = [recall_case("config 1 synthetic (i.i.d. unit rows)", x, q, (10, 50), args.ef_search)]
    g = np.random.default_rng(7)
    docs = g.standard_normal((n // 10, d)).astype(np.float32)
    xc = no.l2_normalize_rows(np.repeat(docs, 10, axis=0) + 0.7 * g.standard_normal((n, d)).astype(np.float32))
    qc = no.l2_normalize_rows(docs[g.choice(n // 10, size=192, replace=False)] +
                              0.7 * g.standard_normal((192, d)).astype(np.float32))
    res.append(recall_case("clustered twin (10 chunks per document)", xc, qc, (10, 50), args.ef_search))
    print(json.dumps({"store": "restated HNSW (oracle/hnsw.c): M=16, ef_construction=100, ef_search=%d, cosine" % args.ef_search,
                      "exact_search_recall": 1.0, "cases": res}))


if __name__ == "__main__":
    main()
