#!/usr/bin/env python
"""BASELINE.json configs[0]: CNIL-corpus-sized synthetic (50k chunks x 1024-d, 48 questions) through the retriever
API: HybridRetriever around the device objects, per question and batched, next to the same retriever logic around the
CPU checkers (exact numpy collection + restated pure-Python rank-bm25 = what the reference runs).  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


class Provider:
    def __init__(self, table):
        self.table = table

    def embed(self, texts):
        return [self.table[t] for t in texts]


class Expander:
    def expand(self, q):
        return [q, q + " reformulation une", q + " reformulation deux", q + " reformulation trois"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=50_000)
    ap.add_argument("--questions", type=int, default=48)
    ap.add_argument("--cpu-questions", type=int, default=3)
    ap.add_argument("--profile", action="store_true", help="cProfile of the per-question device path (top 25 lines on stderr)")
    args = ap.parse_args()
    from b200rag import DeviceCollection, DeviceChunkBM25Index, HybridRetriever, synth, tokenize_french
    import helpers
    from oracle import numpy_oracle as no

    n, d = args.chunks, 1024
    g = np.random.default_rng(1001)
    vocab = [f"mot{i}" for i in range(30_000)]
    p = np.arange(1, len(vocab) + 1, dtype=np.float64) ** (-1.07)
    p /= p.sum()
    lens = g.integers(40, 251, size=n)
    flat = g.choice(len(vocab), size=int(lens.sum()), p=p)
    texts, pos = [], 0
    for ln in lens:
        texts.append(" ".join(vocab[t] for t in flat[pos:pos + ln]))
        pos += ln
    emb = synth.synth_rows(1001, 0, n, d)
    metas = [{"document_path": f"doc_{i // 9}", "chunk_nature": "GUIDE", "chunk_index": i % 9, "confidence": "high",
              "source": "CNIL", "source_url": f"https://www.cnil.fr/fr/doc-{i // 9}"} for i in range(n)]
    ids = [f"doc{i // 9}_{i % 9}" for i in range(n)]
    questions = [" ".join(vocab[t] for t in g.choice(len(vocab), size=10, p=p)) + f" mot{2000 + i} mot{5000 + i}"
                 for i in range(args.questions)]
    qvec = synth.unit_queries(args.questions * 4, d, 2001)
    table = {}
    ex = Expander()
    for i, q in enumerate(questions):
        for j, v in enumerate(ex.expand(q)):
            table[v] = qvec[4 * i + j].tolist()

    t0 = time.perf_counter()
    col = DeviceCollection(dim=d, dtype="f32", capacity=n)
    for s in range(0, n, 5000):
        col.add(ids=ids[s:s + 5000], documents=texts[s:s + 5000], embeddings=emb[s:s + 5000], metadatas=metas[s:s + 5000])
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    bm = DeviceChunkBM25Index()
    bm.build_from_collection(col)
    t_bm25_build = time.perf_counter() - t0
    r = HybridRetriever(collection=col, embedding_provider=Provider(table), chunk_bm25_index=bm, query_expander=ex,
                        enable_summary_prefilter=False)
    r.retrieve_candidates(questions[0], n_candidates=40)
    lat = []
    res_single = []
    for q in questions:
        t0 = time.perf_counter()
        res_single.append(r.retrieve_candidates(q, n_candidates=40))
        lat.append(1e3 * (time.perf_counter() - t0))
    if args.profile:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        for q in questions:
            r.retrieve_candidates(q, n_candidates=40)
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(25)
    r.retrieve_candidates_batch(questions[:4], n_candidates=40)
    t0 = time.perf_counter()
    res_batch = r.retrieve_candidates_batch(questions, n_candidates=40)
    t_batch = time.perf_counter() - t0
    same = all([c.chunk_id for c in a] == [c.chunk_id for c in b] and
               [c.hybrid_score for c in a] == [c.hybrid_score for c in b] for a, b in zip(res_single, res_batch))

    # CPU: same retriever logic around the exact numpy collection + restated pure-Python rank-bm25
    cpu_lat, cpu_same = [], None
    if args.cpu_questions:
        ocol = no.ExactCollection(dim=d)
        ocol._ids, ocol._docs, ocol._metas = list(ids), list(texts), list(metas)
        ocol._x = no.l2_normalize_rows(emb)
        ocol._pos = {i: k for k, i in enumerate(ids)}
        # the oracle's chunked fp64 scoring is slow: use the BASELINE.md Ref-A fp32 BLAS collection for timing
        class FastExact:
            def count(self): return n
            def get(self, **kw): return ocol.get(**kw)
            def query(self, query_embeddings, n_results, where=None, include=None):
                q = no.l2_normalize_rows(np.asarray(query_embeddings, np.float32))
                s = q @ ocol._x.T
                out = {"ids": [], "documents": [], "metadatas": [], "distances": []}
                for b in range(len(q)):
                    idx = np.argpartition(-s[b], n_results - 1)[:n_results]
                    idx = idx[np.lexsort((idx, -s[b][idx]))]
                    out["ids"].append([ids[i] for i in idx]); out["documents"].append([texts[i] for i in idx])
                    out["metadatas"].append([metas[i] for i in idx]); out["distances"].append([float(1 - s[b][i]) for i in idx])
                return out
        t0 = time.perf_counter()
        obm = helpers.OracleChunkBM25Index(tokenize_french)
        obm.build_from_collection(ocol)
        t_cpu_build = time.perf_counter() - t0
        rc = HybridRetriever(collection=FastExact(), embedding_provider=Provider(table), chunk_bm25_index=obm,
                             query_expander=ex, enable_summary_prefilter=False, fuse=helpers.oracle_fuse)
        for q in questions[:args.cpu_questions]:
            t0 = time.perf_counter()
            got = rc.retrieve_candidates(q, n_candidates=40)
            cpu_lat.append(1e3 * (time.perf_counter() - t0))
        cpu_same = [c.chunk_id for c in got] == [c.chunk_id for c in res_single[args.cpu_questions - 1]]
    print(json.dumps({
        "config": f"C1: {n} chunks x {d} fp32, {args.questions} questions, 4 query variants each (dense top-50 + BM25 top-50 "
                  f"per variant, weighted RRF, top-40 candidates) through the retriever API",
        "device_load_s": t_load, "device_bm25_build_s": t_bm25_build,
        "retrieve_candidates_ms_p50": float(np.percentile(lat, 50)), "retrieve_candidates_ms_p99": float(np.percentile(lat, 99)),
        "retrieve_candidates_batch_ms_per_question": 1e3 * t_batch / len(questions),
        "batch_equals_single": bool(same),
        "cpu_reference_logic_ms_per_question": float(np.median(cpu_lat)) if cpu_lat else None,
        "cpu_bm25_build_s": t_cpu_build if cpu_lat else None,
        "cpu_ids_equal_device_ids": cpu_same, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
