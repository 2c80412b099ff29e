#!/usr/bin/env python
"""Auxiliary measurement of BASELINE.json configs[3] (hybrid): 1M chunks, BM25 over CSR postings top-50 + dense
top-50 each, weighted RRF -> top-10.  Not the driver's bench (bench.py is); prints one JSON line.
Parity of every timed result is checked against the oracle on the fly (checker only)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--vocab", type=int, default=200_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--questions", type=int, default=64)
    ap.add_argument("--check", type=int, default=8, help="queries verified against the oracle")
    args = ap.parse_args()
    from b200rag import DeviceCorpus, _lib, rrf_fuse_rows, synth
    from b200rag.bm25 import DeviceBM25, Postings
    from oracle import numpy_oracle as no

    t0 = time.perf_counter()
    docs, n_terms = synth.zipf_corpus(args.docs, args.vocab, seed=1004)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    post = Postings.from_term_ids(docs, n_terms=n_terms)
    t_host_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ix = DeviceBM25(post)
    t_dev_build = time.perf_counter() - t0

    g = np.random.default_rng(2004)
    p = np.arange(1, n_terms + 1, dtype=np.float64) ** (-1.07)
    p /= p.sum()
    Q = args.questions
    # 4 query variants per question (original + 3 expansions), 8-12 Zipf terms + 2 mid-frequency terms
    queries = []
    for _ in range(Q * 4):
        qt = g.choice(n_terms, size=g.integers(8, 13), p=p)
        qt = np.concatenate([qt, g.integers(n_terms // 100, n_terms // 10, size=2)]).astype(np.int32)
        queries.append(qt)
    postings_per_query = float(np.mean([sum(int(post.term_ptr[t + 1] - post.term_ptr[t]) for t in q) for q in queries]))

    # --- BM25 top-50, one query per call (what the reference does) and batched
    for q in queries[:3]:
        ix.search_ids([q], 50)
    lat, dev_ms = [], []
    for q in queries[:64]:
        t0 = time.perf_counter()
        ix.search_ids([q], 50)
        lat.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(float(_lib.last_timings()[0]))
    ix.search_ids(queries, 50)                      # warm-up (first large call allocates scratch)
    t0 = time.perf_counter()
    rows_b, scores_b, counts_b = ix.search_ids(queries, 50)
    t_batch = time.perf_counter() - t0
    dev_batch_ms = float(_lib.last_timings()[0])

    # --- parity against the oracle (numpy CSR restatement of rank-bm25)
    o = no.CsrBM25(docs) if args.check else None
    for i in range(args.check):
        er, es = o.search(queries[i].tolist(), 50)
        assert rows_b[i, :counts_b[i]].tolist() == er.tolist() and np.array_equal(scores_b[i, :counts_b[i]], es), i
    cpu_lat = []
    for i in range(min(args.check, 4)):
        t0 = time.perf_counter()
        o.search(queries[i].tolist(), 50)
        cpu_lat.append(1e3 * (time.perf_counter() - t0))

    # --- dense top-50 for the same questions (4 variants each) on a bf16 corpus
    c = DeviceCorpus(args.dim, "bf16", capacity=args.docs)
    c.fill_synthetic(seed=1004, nrows=args.docs)
    qv = synth.unit_queries(Q * 4, args.dim, 2004)
    c.topk(qv[:8], 50)
    t0 = time.perf_counter()
    rows_d, scores_d, counts_d = c.topk(qv, 50)
    t_dense = time.perf_counter() - t0
    lat_d4 = []
    for i in range(16):
        t0 = time.perf_counter()
        c.topk(qv[4 * i:4 * i + 4], 50)
        lat_d4.append(1e3 * (time.perf_counter() - t0))

    # --- RRF: rankings [dense q0, bm25 q0, dense q1, bm25 q1, ...], reference weights, k=60, top-10 and top-40
    ids = np.full((Q, 8, 50), -1, np.int32)
    for qi in range(Q):
        for v in range(4):
            ids[qi, 2 * v, :counts_d[4 * qi + v]] = rows_d[4 * qi + v, :counts_d[4 * qi + v]]
            ids[qi, 2 * v + 1, :counts_b[4 * qi + v]] = rows_b[4 * qi + v, :counts_b[4 * qi + v]]
    w = np.array([2.0, 3.0, 1.0, 0.75, 1.0, 0.75, 1.0, 0.75])
    rrf_fuse_rows(ids, w, 60, 10)
    t0 = time.perf_counter()
    fi, fs, fc = rrf_fuse_rows(ids, w, 60, 10)
    t_rrf = time.perf_counter() - t0
    from oracle import c_oracle
    for qi in range(min(Q, args.check)):
        ei, es = c_oracle.rrf(ids[qi], w, 60, 10)
        assert fi[qi, :fc[qi]].tolist() == ei.tolist() and np.array_equal(fs[qi, :fc[qi]], es)

    print(json.dumps({
        "config": f"hybrid: {args.docs} chunks, vocab {n_terms}, doc len U[40,250], Zipf 1.07; {Q} questions x 4 query variants",
        "postings_nnz": int(len(post.post_row)), "avg_postings_per_query": postings_per_query,
        "host_corpus_gen_s": t_gen, "host_csr_build_s": t_host_build, "device_build_s": t_dev_build,
        "bm25_top50_single_ms_p50": float(np.percentile(lat, 50)), "bm25_top50_single_ms_p99": float(np.percentile(lat, 99)),
        "bm25_top50_single_device_ms_p50": float(np.percentile(dev_ms, 50)),
        "bm25_top50_batch_queries_per_s": len(queries) / t_batch, "bm25_batch_device_ms": dev_batch_ms,
        "bm25_algorithmic_GBps_batch_device": postings_per_query * 12 * len(queries) / (dev_batch_ms / 1e3) / 1e9,
        "bm25_algorithmic_GBps_single": postings_per_query * 12 / (np.percentile(lat, 50) / 1e3) / 1e9,
        "cpu_oracle_csr_numpy_ms": float(np.median(cpu_lat)) if cpu_lat else None,
        "dense_top50_bf16_queries_per_s": len(qv) / t_dense, "dense_top50_4query_call_ms_p50": float(np.percentile(lat_d4, 50)),
        "rrf_questions_per_s": Q / t_rrf, "rrf_batch_ms": 1e3 * t_rrf,
        "parity_checked_queries": args.check, "counters": _lib.counters()}))


if __name__ == "__main__":
    main()
