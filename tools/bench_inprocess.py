#!/usr/bin/env python
"""ONE process driving all GPUs of the box behind the drop-in (the reference is a single Streamlit process,
app.py:42-119): a sharded DeviceCorpus (block-cyclic rows over the shard slots, per-shard exact top-k written into the
primary GPU's gather buffer over NVLink peer memory, merged there) at the config-5 shape, and a sharded BM25 index at
the config-4 shape.  Every timed result is checked against the oracle (checker only).  Prints one JSON line.
   python tools/bench_inprocess.py --shards 8"""
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
    ap.add_argument("--shards", type=int, default=2)
    ap.add_argument("--rows-per-shard", type=int, default=12_500_000)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--bm25-docs", type=int, default=1_000_000)
    args = ap.parse_args()
    import bench
    from b200rag import DeviceCorpus, _lib, synth
    from b200rag.bm25 import DeviceBM25, Postings
    from oracle import c_oracle

    G, d, k = args.shards, 1024, args.k
    n = G * args.rows_per_shard
    out = {"shards": G, "rows": n, "dtype": args.dtype, "k": k}
    t0 = time.perf_counter()
    c = DeviceCorpus(d, args.dtype, capacity=n, n_shards=G)
    step = 25_000_000
    for r0 in range(0, n, step):
        c.fill_synthetic(seed=1005, nrows=min(step, n - r0))
    out["fill_s"] = time.perf_counter() - t0
    q = synth.unit_queries(max(args.batch, 64), d, 2005)
    for _ in range(3):
        c.topk(q[:1], k)
    lat = []
    for i in range(40):
        t0 = time.perf_counter()
        c.topk(q[i % 32:i % 32 + 1], k)
        lat.append(1e3 * (time.perf_counter() - t0))
    out["batch1_call_ms_p50"], out["batch1_call_ms_p99"] = float(np.percentile(lat, 50)), float(np.percentile(lat, 99))
    row_bytes = d * (4 if args.dtype == "f32" else 2)
    out["batch1_aggregate_GBps"] = n * row_bytes / (out["batch1_call_ms_p50"] / 1e3) / 1e9
    c.topk(q[:args.batch], k)
    tb = []
    for _ in range(3):
        t0 = time.perf_counter()
        rows, scores, counts = c.topk(q[:args.batch], k)
        tb.append(time.perf_counter() - t0)
    out["batch_call_ms"] = 1e3 * float(np.median(tb))
    out["batch_queries_per_s"] = args.batch / float(np.median(tb))
    # parity: 64 queries re-scored by oracle.c + completeness on two row slices
    r64, s64, c64 = c.topk(q[:64], k)
    assert np.array_equal(r64, rows[:64]) and np.array_equal(s64, scores[:64])
    slices = [(n - 300_000, n - 200_000), (n // 3, n // 3 + 50_000)]
    pairs = bench.oracle_check_slices(c, q[:64], r64.astype(np.int64), s64, args.dtype, slices)
    out["dense_parity"] = f"ok: 64 queries, {pairs} (query, row) pairs against oracle.c"
    c.close()

    # ---- BM25 sharded over the same slots
    doc_ptr, tokens, n_terms = bench.zipf_tokens(args.bm25_docs, 200_000, 1004)
    post = Postings.from_flat_tokens(doc_ptr, tokens, n_terms)
    t0 = time.perf_counter()
    ix = DeviceBM25(post, n_shards=G)
    out["bm25_build_s"] = time.perf_counter() - t0
    g = np.random.default_rng(2004)
    p = np.arange(1, n_terms + 1, dtype=np.float64) ** (-1.07)
    p /= p.sum()
    queries = [np.concatenate([g.choice(n_terms, size=g.integers(8, 13), p=p),
                               g.integers(n_terms // 100, n_terms // 10, size=2)]).astype(np.int32) for _ in range(256)]
    ix.search_ids(queries, 50)
    tb = []
    for _ in range(3):
        t0 = time.perf_counter()
        rb, sb, cb = ix.search_ids(queries, 50)
        tb.append(time.perf_counter() - t0)
    out["bm25_batch_queries_per_s"] = len(queries) / float(np.median(tb))
    lat = []
    for i in range(30):
        t0 = time.perf_counter()
        ix.search_ids([queries[i]], 50)
        lat.append(1e3 * (time.perf_counter() - t0))
    out["bm25_single_call_ms_p50"] = float(np.percentile(lat, 50))
    for i in range(8):
        want = c_oracle.bm25_scores(post.term_ptr, post.post_row, post.post_tf, post.doc_len, post.idf, post.avgdl,
                                    post.k1, post.b, queries[i])
        er, es = c_oracle.bm25_select(want, 50)
        assert rb[i, :cb[i]].tolist() == er.tolist() and np.array_equal(sb[i, :cb[i]], es), i
    out["bm25_parity"] = "ok: 8 queries bit-equal to oracle.c (global idf / avgdl on every shard)"
    ix.close()
    out["counters"] = _lib.counters()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
