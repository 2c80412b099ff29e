"""GPU parity tests (run with -m gpu on a B200): every call goes through the
C ABI (libb200rag.so) and is compared bit-for-bit with the oracle / the golden
vectors produced by the reference's own code."""
import os
import sys

import numpy as np
import pytest

import helpers
from conftest import load_golden, unhex
from oracle import c_oracle
from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu

DT = {"f32": no.DT_F32, "bf16": no.DT_BF16, "f16": no.DT_F16}


def raw_rows(x32, dt):
    """storage-dtype bit patterns for the C oracle, from fp32 values that are already representable"""
    if dt == no.DT_F32:
        return np.ascontiguousarray(x32, np.float32)
    if dt == no.DT_BF16:
        return no.f32_to_bf16_bits(x32)
    return x32.astype(np.float16).view(np.uint16)


def check_topk(corpus, q, k, dt, allow=None):
    bitmap = np.packbits(allow, bitorder="little") if allow is not None else None
    rows, scores, counts = corpus.topk(q, k, bitmap)
    stored = corpus.download()
    er, es, ec = c_oracle.dense_topk(q, raw_rows(stored, dt), dt, k, bitmap)
    assert counts.tolist() == ec.tolist()
    for b in range(len(q)):
        c = counts[b]
        assert rows[b, :c].tolist() == er[b, :c].tolist(), f"query {b}"
        assert [float(v).hex() for v in scores[b, :c]] == [float(v).hex() for v in es[b, :c]]
        assert (rows[b, c:] == -1).all()
    return rows, scores, counts


def test_device_is_b200():
    from b200rag import _lib
    info = _lib.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_dense_golden(dense_small, dtype):
    from b200rag import DeviceCorpus
    x, q, gold = dense_small
    c = DeviceCorpus(x.shape[1], dtype)
    c.append(x)
    assert np.array_equal(c.download(), no.quantize(x, DT[dtype]))      # RNE conversion on the device
    allow = (np.arange(len(x)) % 3 != 0)
    bitmap = np.packbits(allow, bitorder="little")
    for case in gold["cases"]:
        if case["dtype"] != dtype:
            continue
        rows, scores, counts = c.topk(q, case["k"], bitmap if case["filtered"] else None)
        for qi in range(len(q)):
            assert rows[qi, :counts[qi]].tolist() == case["rows"][qi], (case["k"], case["filtered"], qi)
            assert [float(v).hex() for v in scores[qi, :counts[qi]]] == case["scores"][qi]
    c.close()


@pytest.mark.parametrize("dtype,n,d", [("bf16", 20011, 1024), ("f32", 9001, 1024), ("f16", 5000, 768),
                                       ("bf16", 3000, 64), ("f32", 777, 256), ("bf16", 4097, 2048)])
def test_dense_random_vs_oracle(dtype, n, d):
    from b200rag import DeviceCorpus
    x = helpers.synth_unit(n, d, seed=n)
    q = helpers.synth_unit(7, d, seed=n + 1)
    q[2] = x[n // 2]                                   # a query equal to a row
    c = DeviceCorpus(d, dtype)
    c.append(x[: n // 3])
    c.append(x[n // 3:])                               # appended in two uploads
    assert c.count() == n
    for B, k in [(1, 10), (2, 50), (3, 1), (4, 100), (7, 10), (5, 224)]:
        check_topk(c, q[:B], min(k, n), DT[dtype])
    allow = np.random.default_rng(5).random(n) < 0.3
    check_topk(c, q[:3], 10, DT[dtype], allow)
    allow[:] = False
    allow[[3, n - 1]] = True
    rows, scores, counts = check_topk(c, q[:2], 10, DT[dtype], allow)
    assert counts.tolist() == [2, 2]
    c.close()


def test_dense_many_exact_ties_take_the_fallback_pass():
    """More identical best rows than the candidate list holds: the margin check must fail, the
    fallback pass must run, and ties must resolve to the lowest rows."""
    from b200rag import DeviceCorpus, _lib
    n, d = 6000, 128
    x = helpers.synth_unit(n, d, seed=2)
    q = helpers.synth_unit(3, d, seed=3)
    dup = np.random.default_rng(9).choice(n, size=700, replace=False)
    x[dup] = q[0]                                      # 700 rows tie for the best score of query 0
    c = DeviceCorpus(d, "f32")
    c.append(x)
    before = _lib.counters()["fallbacks"]
    rows, scores, counts = check_topk(c, q, 10, no.DT_F32)
    assert rows[0].tolist() == sorted(dup.tolist())[:10]
    assert _lib.counters()["fallbacks"] == before + 1
    check_topk(c, q[:1], 100, no.DT_F32)
    c.close()


def test_dense_near_ties_within_filter_error():
    """rows that differ from the best row by less than the fp32 filter error: order must still be the fp64 order"""
    from b200rag import DeviceCorpus
    n, d = 4000, 1024
    g = np.random.default_rng(11)
    x = helpers.synth_unit(n, d, seed=12)
    q = helpers.synth_unit(1, d, seed=13)
    for j in range(60):                                # 60 tiny perturbations of the query itself
        x[100 + 37 * j] = no.l2_normalize_rows(q[0] + 3e-5 * g.standard_normal(d).astype(np.float32))[0]
    for dtype in ("f32", "bf16"):
        c = DeviceCorpus(d, dtype)
        c.append(x)
        check_topk(c, q, 10, DT[dtype])
        check_topk(c, q, 50, DT[dtype])
        c.close()


@pytest.mark.parametrize("dtype,dim", [("bf16", 1024), ("f16", 1024), ("bf16", 2048), ("f32", 1024), ("bf16", 256)])
def test_tensor_core_filter_error_stays_inside_the_bound(dtype, dim, record_property):
    """The margin check is only as sound as eps, the bound on |tensor-core filter score - exact score|
    (accumulation part: eps_rel * |q| * max|x|).  Adversarial operands, all exactly representable in the 16-bit
    operand format so that accumulation is the ONLY error: (a) aligned all-positive products (the accumulator is
    as large as it gets, every truncation goes the same way), (b) alternating-sign pairs inside every K=16 block
    (the sum cancels to ~0 while sum|products| stays |q||x|), (c) 2^-10..1 dynamic range inside every block with
    random signs, (d) ordinary unit vectors.  > 1e5 (query, row) pairs per case; the observed max error / bound is
    recorded and must stay below 1/2."""
    from b200rag import DeviceCorpus
    n, B = 4096, 256
    g = np.random.default_rng(dim + len(dtype))
    dt = DT[dtype]
    sig = 8 if dtype != "f16" else 11                 # significand bits of the operand format

    def representable(a):                              # round to `sig` bits, exponents inside fp16's normal range
        a = np.asarray(a, np.float64)
        e = np.floor(np.log2(np.maximum(np.abs(a), 2.0 ** -13)))
        return (np.round(a / 2.0 ** (e - sig + 1)) * 2.0 ** (e - sig + 1)).astype(np.float32)

    mag = representable(2.0 ** g.uniform(-3, 0, size=(n, dim)))
    cases = {}
    cases["aligned_positive"] = (mag, representable(mag[g.integers(0, n, B)] * 2.0 ** g.integers(-1, 2, size=(B, 1))))
    alt = np.where(np.arange(dim) % 2 == 0, 1.0, -1.0).astype(np.float32)
    xa = mag.copy()
    xa[:, 1::2] = xa[:, 0::2]                          # pairs of equal magnitude ...
    cases["alternating_cancel"] = (xa * alt, representable(np.abs(xa[g.integers(0, n, B)]) *
                                                           (1 + 2.0 ** -6 * g.integers(0, 3, size=(B, dim)))))
    wide = representable(2.0 ** g.uniform(-10, 0, size=(n, dim)) * g.choice([-1.0, 1.0], size=(n, dim)))
    cases["wide_range"] = (wide, representable(2.0 ** g.uniform(-10, 0, size=(B, dim)) * g.choice([-1.0, 1.0], size=(B, dim))))
    cases["unit"] = (representable(helpers.synth_unit(n, dim, seed=5)), representable(helpers.synth_unit(B, dim, seed=6)))
    worst = 0.0
    for name, (x, q) in cases.items():
        c = DeviceCorpus(dim, dtype)
        c.append(x)
        stored = c.download()
        if dtype != "f32":
            assert np.array_equal(stored, x), name   # representable: stored as given
        rows, scores, counts = c.topk(q, 10)
        er, es, ec = c_oracle.dense_topk(q, raw_rows(stored, dt), dt, 10)
        assert rows.tolist() == er.tolist() and np.array_equal(scores, es), name
        cr, cs, cc, eps_rel = c.debug_last_candidates(B, 8192)
        assert (cc >= 0).all(), name
        # operands of the contraction: the stored rows (their bf16 shadow for an fp32 corpus) and the rounded queries
        xo = stored if dtype != "f32" else no.quantize(stored, no.DT_BF16)
        qo = no.quantize(q, no.DT_F16 if dtype == "f16" else no.DT_BF16)
        assert np.array_equal(qo, q), name
        exact = qo.astype(np.float64) @ xo.astype(np.float64).T            # fp64: products exact, sum to ~1e-16
        bound = eps_rel * np.linalg.norm(qo.astype(np.float64), axis=1) * np.linalg.norm(xo.astype(np.float64), axis=1).max()
        n_pairs, ratio = 0, 0.0
        for b in range(B):
            r = cr[b, :cc[b]]
            err = np.abs(cs[b, :cc[b]].astype(np.float64) - exact[b, r])
            ratio = max(ratio, float(err.max() / bound[b]) if len(r) else 0.0)
            n_pairs += len(r)
        assert n_pairs >= 100_000, (name, n_pairs)
        record_property(f"{name}_max_error_over_bound", ratio)
        print(f"[eps] {dtype} dim={dim} {name}: {n_pairs} pairs, max |filter - exact| / bound = {ratio:.4f}")
        worst = max(worst, ratio)
        c.close()
    assert worst <= 0.5, worst


def test_dense_edge_cases():
    from b200rag import DeviceCorpus, DeviceCollection, B200RagError
    d = 128
    c = DeviceCorpus(d, "bf16")
    q = helpers.synth_unit(2, d, seed=1)
    rows, scores, counts = c.topk(q, 5)                 # empty corpus
    assert counts.tolist() == [0, 0] and (rows == -1).all()
    x = helpers.synth_unit(3, d, seed=2)
    c.append(x)
    rows, scores, counts = check_topk(c, q, 10, no.DT_BF16)   # fewer rows than k
    assert counts.tolist() == [3, 3]
    # k above the fused select's limit is served in several passes (still exact), not refused
    rows, scores, counts = c.topk(q, 225)
    assert counts.tolist() == [3, 3] and rows.shape == (2, 225)
    with pytest.raises(B200RagError):
        DeviceCorpus(100, "bf16")
    with pytest.raises(ValueError):
        c.topk(np.zeros((1, 64), np.float32), 1)
    c.close()
    col = DeviceCollection(dim=d, dtype="f32")
    out = col.query(query_embeddings=[q[0].tolist()], n_results=5)
    assert out["ids"] == [[]] and out["distances"] == [[]]


def test_collection_matches_exact_collection(e2e_data):
    from b200rag import DeviceCollection
    gold, emb, table = e2e_data
    for dtype in ("f32", "bf16"):
        dev = DeviceCollection(dim=emb.shape[1], dtype=dtype)
        ora = no.ExactCollection(dim=emb.shape[1], dtype=DT[dtype])
        helpers.fill(dev, gold["chunks"], emb)
        helpers.fill(ora, gold["chunks"], emb)
        assert dev.count() == ora.count()
        qs = [list(map(float, v)) for v in list(table.values())[:6]]
        wheres = [None, {"source": "CNIL"}, {"source": {"$ne": "ENTREPRISE"}}, {"tag_rh": True},
                  {"$or": [{"source": "CNIL"}, {"$and": [{"source": "ENTREPRISE"}, {"tag_rh": True}]}]},
                  {"chunk_nature": {"$in": ["GUIDE", "SANCTION"]}}, {"source": "NOPE"}]
        for w in wheres:
            for n_res in (50, 7):
                assert dev.query(query_embeddings=qs[:2], n_results=n_res, where=w,
                                 include=["documents", "metadatas", "distances"]) == \
                    ora.query(query_embeddings=qs[:2], n_results=n_res, where=w,
                              include=["documents", "metadatas", "distances"])
        a = dev.get(limit=20, offset=10, include=["documents", "metadatas", "embeddings"])
        b = ora.get(limit=20, offset=10, include=["documents", "metadatas", "embeddings"])
        assert a["ids"] == b["ids"] and a["documents"] == b["documents"] and a["metadatas"] == b["metadatas"]
        assert np.array_equal(a["embeddings"], b["embeddings"])
        # delete + update, then the same queries again
        victims = [c["id"] for c in gold["chunks"][5:60:3]]
        dev.delete(ids=victims); ora.delete(ids=victims)
        dev.delete(where={"tag_rh": True}); ora.delete(where={"tag_rh": True})
        some = dev.get(limit=3)["ids"]
        dev.update(ids=some, metadatas=[{"source": "CNIL", "document_path": "x"}] * 3)
        ora.update(ids=some, metadatas=[{"source": "CNIL", "document_path": "x"}] * 3)
        assert dev.count() == ora.count()
        assert dev.query(query_embeddings=qs[2:4], n_results=40, where={"source": "CNIL"}) == \
            ora.query(query_embeddings=qs[2:4], n_results=40, where={"source": "CNIL"})
        assert np.array_equal(dev.get(include=["embeddings"])["embeddings"], ora.get(include=["embeddings"])["embeddings"])


def test_synthetic_fill_matches_numpy_twin():
    from b200rag import DeviceCorpus
    from b200rag import synth
    for dtype in ("f32", "bf16"):
        c = DeviceCorpus(256, dtype)
        c.fill_synthetic(seed=77, nrows=1000)
        c.fill_synthetic(seed=77, nrows=500, gen_row0=5000)
        want = np.concatenate([synth.synth_rows(77, 0, 1000, 256), synth.synth_rows(77, 5000, 500, 256)])
        assert np.array_equal(c.download(), no.quantize(want, DT[dtype]))
        q = synth.unit_queries(2, 256, 1)
        check_topk(c, q, 10, DT[dtype])
        c.close()


def test_collection_save_load_roundtrip(e2e_data, tmp_path):
    from b200rag import DeviceCollection
    gold, emb, table = e2e_data
    for dtype in ("bf16", "f32"):
        a = DeviceCollection(dim=emb.shape[1], dtype=dtype)
        helpers.fill(a, gold["chunks"], emb)
        a.save(str(tmp_path / dtype))
        b = DeviceCollection.load(str(tmp_path / dtype))
        assert b.count() == a.count() and np.array_equal(a.corpus.download(), b.corpus.download())
        qs = [list(map(float, v)) for v in list(table.values())[:3]]
        assert a.query(query_embeddings=qs, n_results=20, where={"source": "CNIL"}) == \
            b.query(query_embeddings=qs, n_results=20, where={"source": "CNIL"})


def test_selective_where_filters_through_the_tensor_core_path():
    from b200rag import DeviceCorpus
    n, d = 300_000, 256
    c = DeviceCorpus(d, "bf16")
    c.fill_synthetic(seed=9, nrows=n)
    q = helpers.synth_unit(9, d, seed=10)
    g = np.random.default_rng(0)
    for frac in (0.0005, 0.02, 0.5):
        allow = g.random(n) < frac
        check_topk(c, q, 10, no.DT_BF16, allow)
    allow = np.zeros(n, bool)
    allow[250_000:250_300] = True                       # a contiguous island of allowed rows
    check_topk(c, q, 50, no.DT_BF16, allow)
    c.close()


# ---------------------------------------------------------------- BM25 ----
def test_bm25_golden_from_reference_index():
    from b200rag import DeviceCollection, DeviceChunkBM25Index
    gold = load_golden("bm25_small.json")
    col = DeviceCollection(dim=64, dtype="f32")
    emb = helpers.synth_unit(len(gold["chunks"]), 64, seed=3)
    helpers.fill(col, gold["chunks"], emb)
    idx = DeviceChunkBM25Index()
    with pytest.raises(RuntimeError):
        idx.search("x")
    idx.build_from_collection(col, batch_size=70)
    assert idx.is_built and idx.chunk_ids == gold["kept_ids"]
    for case in gold["cases"]:
        res = idx.search(case["query"], top_k=case["top_k"],
                         doc_filter=set(case["doc_filter"]) if case["doc_filter"] is not None else None)
        assert [{"doc_key": r.doc_key, "score": float(r.score).hex()} for r in res] == case["results"], case["query"]
        for r in res:
            assert "text" in r.metadata and r.metadata["document_path"]
    col2 = DeviceCollection(dim=64, dtype="f32")
    helpers.fill(col2, gold["common_corpus"], helpers.synth_unit(6, 64, seed=1))
    idx2 = DeviceChunkBM25Index()
    idx2.build_from_collection(col2)
    assert idx2.search("données rgpd", top_k=10) == []


def test_chunk_bm25_incremental_add_remove_equals_fresh_build():
    """DeviceChunkBM25Index.add_chunks / remove_chunks (enterprise ingestion adds and deletes chunks,
    src/processing/ingest_enterprise.py:241-309) == build_from_collection over the collection in the same state:
    vocabulary order, idf bits, results.  The golden cases of the reference's own index still hold after growing
    the index chunk by chunk."""
    from b200rag import DeviceCollection, DeviceChunkBM25Index
    gold = load_golden("bm25_small.json")
    chunks = gold["chunks"]
    emb = helpers.synth_unit(len(chunks), 64, seed=3)
    n0 = len(chunks) * 2 // 3
    col = DeviceCollection(dim=64, dtype="f32")
    helpers.fill(col, chunks[:n0], emb[:n0])
    idx = DeviceChunkBM25Index()
    idx.build_from_collection(col)
    added = idx.add_chunks([c["id"] for c in chunks[n0:]], [c["text"] for c in chunks[n0:]],
                           [c["metadata"] for c in chunks[n0:]])
    first = {c["id"] for c in chunks[:n0]}
    assert idx.chunk_ids == gold["kept_ids"] and added == sum(1 for i in gold["kept_ids"] if i not in first)
    for case in gold["cases"]:
        res = idx.search(case["query"], top_k=case["top_k"],
                         doc_filter=set(case["doc_filter"]) if case["doc_filter"] is not None else None)
        assert [{"doc_key": r.doc_key, "score": float(r.score).hex()} for r in res] == case["results"], case["query"]
    # remove a third of the chunks: equal to a fresh build over what remains
    gone = gold["kept_ids"][::3]
    assert idx.remove_chunks(gone) == len(gone)
    col2 = DeviceCollection(dim=64, dtype="f32")
    left = [c for c in chunks if c["id"] not in set(gone)]
    helpers.fill(col2, left, helpers.synth_unit(len(left), 64, seed=4))
    fresh = DeviceChunkBM25Index()
    fresh.build_from_collection(col2)
    assert idx.chunk_ids == fresh.chunk_ids and idx.postings.vocab == fresh.postings.vocab
    assert np.array_equal(idx.postings.idf, fresh.postings.idf) and np.array_equal(idx.postings.post_row, fresh.postings.post_row)
    for case in gold["cases"]:
        a = idx.search(case["query"], top_k=case["top_k"])
        b = fresh.search(case["query"], top_k=case["top_k"])
        assert [(r.doc_key, float(r.score).hex()) for r in a] == [(r.doc_key, float(r.score).hex()) for r in b]
    assert idx.add_chunks(["empty"], ["   "], [{}]) == 0


def test_summary_bm25_golden(golden_dir):
    from b200rag import DeviceSummaryBM25Index
    gold = load_golden("summary_bm25.json")
    sm = DeviceSummaryBM25Index()
    sm.build(os.path.join(golden_dir, "summaries_input.json"))
    assert sm._is_built and sm.doc_keys == gold["doc_keys"]
    for case in gold["cases"]:
        res = sm.search(case["query"], top_k=case["top_k"])
        assert [{"doc_key": r.doc_key, "score": float(r.score).hex()} for r in res] == case["results"]


@pytest.mark.parametrize("n_docs,vocab", [(5000, 2000), (60000, 30000)])
def test_bm25_scores_and_select_vs_oracle(n_docs, vocab):
    from b200rag.bm25 import DeviceBM25, Postings
    docs, n_terms = helpers.zipf_docs(n_docs, vocab, seed=n_docs)
    p = Postings.from_term_ids(docs, n_terms=n_terms)
    o = no.CsrBM25(docs)
    assert np.array_equal(o.idf, p.idf) and np.array_equal(o.post_row, p.post_row)
    ix = DeviceBM25(p)
    g = np.random.default_rng(1)
    allow = g.random(n_docs) < 0.5
    bitmap = np.packbits(allow, bitorder="little")
    queries = []
    for _ in range(12):
        qt = g.integers(0, min(n_terms, 400), size=g.integers(1, 14)).astype(np.int32)
        if g.random() < 0.4:
            qt[g.integers(0, len(qt))] = -1               # unknown token
        if g.random() < 0.4:
            qt = np.concatenate([qt, qt[:2]])             # repeated tokens are added again
        queries.append(qt)
    for qt in queries[:4]:
        want = c_oracle.bm25_scores(o.term_ptr, o.post_row, o.post_tf, o.doc_len.astype(np.int32), o.idf, o.avgdl,
                                    1.5, 0.75, qt)
        got = ix.scores(qt)
        assert np.array_equal(got, want)
        assert np.array_equal(want, o.get_scores(qt.tolist()))
    for k in (50, 10, 224):
        rows, scores, counts = ix.search_ids(queries, k)
        rows_f, scores_f, counts_f = ix.search_ids(queries, k, bitmap)
        for i, qt in enumerate(queries):
            er, es = o.search(qt.tolist(), k)
            assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es)
            er, es = o.search(qt.tolist(), k, allow)
            assert rows_f[i, :counts_f[i]].tolist() == er.tolist() and np.array_equal(scores_f[i, :counts_f[i]], es)
    # accumulator is clean after all of that
    assert not ix.scores(np.array([-1], np.int32)).any()
    ix.close()


# ----------------------------------------------------------------- RRF ----
def test_rrf_golden_from_reference():
    from b200rag import reciprocal_rank_fusion, fuse_ranked
    for case in load_golden("rrf.json"):
        kw = {"k": case["k"]}
        if case["weights"] is not None:
            kw["weights"] = case["weights"]
        got = reciprocal_rank_fusion(case["rankings"], **kw)
        assert {i: float(s).hex() for i, s in got.items()} == case["scores"]
        assert list(got.keys()) == list(case["scores"].keys())            # dict in first-seen order
        ids, sc = fuse_ranked(case["rankings"], **kw)
        assert ids == case["order"]
        ids10, _ = fuse_ranked(case["rankings"], top=10, **kw)
        assert ids10 == case["order"][:10]


def test_rrf_batched_vs_oracle():
    from b200rag import rrf_fuse_rows
    g = np.random.default_rng(3)
    Q, R, L = 64, 8, 50
    ids = np.full((Q, R, L), -1, np.int32)
    for qi in range(Q):
        for r in range(R):
            ln = g.integers(0, L + 1)
            ids[qi, r, :ln] = g.choice(400, size=ln, replace=False)
    w = np.array([2.0, 3.0, 1.0, 0.75, 1.0, 0.75, 1.0, 0.75])
    out_ids, out_sc, counts = rrf_fuse_rows(ids, w, 60, 40)
    for qi in range(Q):
        ei, es = c_oracle.rrf(ids[qi], w, 60, 40)
        assert out_ids[qi, :counts[qi]].tolist() == ei.tolist()
        assert np.array_equal(out_sc[qi, :counts[qi]], es)


# ----------------------------------------------------------------- e2e ----
def test_hybrid_retrieval_end_to_end_matches_reference(e2e_data, golden_dir):
    """HybridRetriever around DeviceCollection + DeviceChunkBM25Index + DeviceSummaryBM25Index + device RRF
    returns exactly what the reference's RAGRetriever returned around the exact CPU collection."""
    from b200rag import DeviceCollection, DeviceChunkBM25Index, DeviceSummaryBM25Index, HybridRetriever
    gold, emb, table = e2e_data
    col = DeviceCollection(dim=emb.shape[1], dtype="f32")
    helpers.fill(col, gold["chunks"], emb)
    bm = DeviceChunkBM25Index()
    bm.build_from_collection(col)
    sm = DeviceSummaryBM25Index()
    sm.build(os.path.join(golden_dir, "e2e_summaries.json"))
    for run in gold["runs"]:
        cands, docs = helpers.run_e2e_case(HybridRetriever, col, bm, sm, gold, table, run)
        assert cands == run["candidates"], (run["config"], run["query"])
        assert docs == run["documents"], (run["config"], run["query"])


# ------------------------------------------------------- multi-GPU tail ----
def test_merge_topk_dev_vs_host_merge():
    import torch
    from b200rag import _lib
    from b200rag.sharded import merge_lists_torch
    g = torch.Generator().manual_seed(0)
    for G, B, k in [(2, 5, 10), (8, 33, 100), (3, 1, 1), (8, 4, 224)]:
        scores = torch.randn(G, B, k, generator=g, dtype=torch.float64).sort(dim=2, descending=True).values
        if k > 1:
            scores[:, :, 1] = scores[:, :, 0]                          # ties inside and across lists
        scores[1:] = torch.where(torch.rand(G - 1, B, k, generator=g) < 0.2, scores[:1], scores[1:])
        ids = torch.stack([torch.randperm(100000, generator=g)[: B * k].reshape(B, k) + 100000 * r for r in range(G)])
        ids[-1, :, k // 2:] = -1                                        # a short list
        want_s, want_i, want_c = merge_lists_torch(scores, ids, k)
        ds, di = scores.cuda(), ids.cuda()
        o_s = torch.empty(B, k, dtype=torch.float64, device="cuda")
        o_i = torch.empty(B, k, dtype=torch.int64, device="cuda")
        o_c = torch.empty(B, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        _lib.check(_lib.lib().rag_merge_topk_dev(ds.data_ptr(), di.data_ptr(), G, B, k, 0, o_s.data_ptr(),
                                                 o_i.data_ptr(), o_c.data_ptr()))
        torch.cuda.synchronize()
        # packed layout: one buffer per rank holding [scores | ids], as bench.py all-gathers it
        packed = torch.empty((G, 2, B, k), dtype=torch.float64, device="cuda")
        packed[:, 0] = ds
        packed[:, 1].view(torch.int64).copy_(di)
        o_s2, o_i2, o_c2 = torch.empty_like(o_s), torch.empty_like(o_i), torch.empty_like(o_c)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().rag_merge_topk_dev(packed.data_ptr(), packed.data_ptr() + B * k * 8, G, B, k, 2 * B * k,
                                                 o_s2.data_ptr(), o_i2.data_ptr(), o_c2.data_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(o_i2, o_i) and torch.equal(o_s2, o_s) and torch.equal(o_c2, o_c)
        assert torch.equal(o_i.cpu(), want_i) and torch.equal(o_s.cpu(), want_s) and torch.equal(o_c.cpu(), want_c)


def test_peer_exchange_single_rank_roundtrip():
    """rag_exchange_* with world = 1: the rank pushes its block into its own buffer, the merge kernel waits for the
    epoch flag and merges.  Several epochs (both buffer parities, flag reuse); the multi-rank path runs under
    torchrun (bench.py --gpus N, tests/test_dist.py covers the host logic on gloo)."""
    import ctypes as C
    import torch
    from b200rag import _lib
    from b200rag.sharded import merge_lists_torch
    L = _lib.lib()
    h = C.c_void_p()
    handle = (C.c_uint8 * 64)()
    _lib.check(L.rag_exchange_create(C.byref(h), 1, 0, 1 << 20, handle))
    _lib.check(L.rag_exchange_connect(h, handle))
    g = torch.Generator().manual_seed(1)
    try:
        for epoch, (B, k) in enumerate([(7, 10), (1, 1), (64, 100), (7, 10), (300, 50)]):
            scores = torch.randn(1, B, k, generator=g, dtype=torch.float64).sort(dim=2, descending=True).values
            ids = torch.randperm(10 ** 6, generator=g)[: B * k].reshape(1, B, k)
            if k > 2:
                ids[0, :, -1] = -1
            want_s, want_i, want_c = merge_lists_torch(scores, ids, k)
            ds, di = scores[0].contiguous().cuda(), ids[0].contiguous().cuda()
            o_s = torch.empty(B, k, dtype=torch.float64, device="cuda")
            o_i = torch.empty(B, k, dtype=torch.int64, device="cuda")
            o_c = torch.empty(B, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            _lib.check(L.rag_exchange_merge_topk_dev(h, ds.data_ptr(), di.data_ptr(), B, k, o_s.data_ptr(),
                                                     o_i.data_ptr(), o_c.data_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(o_i.cpu(), want_i) and torch.equal(o_s.cpu(), want_s) and torch.equal(o_c.cpu(), want_c)
        # a payload larger than the slot is refused, not truncated
        assert L.rag_exchange_merge_topk_dev(h, ds.data_ptr(), di.data_ptr(), 4096, 224, o_s.data_ptr(),
                                             o_i.data_ptr(), o_c.data_ptr()) != 0
    finally:
        _lib.check(L.rag_exchange_destroy(h))


def test_two_rank_peer_exchange_vs_oracle():
    """one process per GPU (torch.distributed.run, 2 ranks): row shards in DeviceCorpus, the stream-ordered device
    step with the NVLink peer-memory exchange (different B / k per case, a delayed rank, several epochs) and the
    host-buffer call, merged result compared with the oracle over the whole corpus on every rank
    (tests/dist_gpu_worker.py).  Needs 2 GPUs: skipped on a 1-GPU box."""
    import socket
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "dist_gpu_worker.py")]
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout[-3000:], res.stderr[-3000:])


# --------------------------------------------------- BASELINE-size cases ----
def test_dense_config2_full_size_properties():
    """BASELINE config 2 (1M x 1024 fp32, top-10): size-independent properties + two queries against the oracle."""
    from b200rag import DeviceCorpus, synth
    n, d = 1_000_000, 1024
    c = DeviceCorpus(d, "f32", capacity=n)
    c.fill_synthetic(seed=1002, nrows=n)
    q = synth.unit_queries(4, d, 2002)
    planted = [123456, 999999, 0]
    rows_p = np.concatenate([c.download(r, 1) for r in planted])
    qq = np.concatenate([q, rows_p])
    rows, scores, counts = c.topk(qq, 10)
    assert (counts == 10).all()
    assert (np.diff(scores, axis=1) <= 0).all()                        # sorted
    for j, r in enumerate(planted):                                    # a row is its own nearest neighbour
        assert rows[4 + j, 0] == r and abs(scores[4 + j, 0] - 1.0) < 1e-6
    r1, s1, _ = c.topk(qq[:1], 10)                                     # batch-1 == batch-7 row
    assert r1[0].tolist() == rows[0].tolist() and np.array_equal(s1[0], scores[0])
    # against the oracle on the regenerated rows (numpy twin of the generator), 2 queries over all rows ...
    x = synth.synth_rows(1002, 0, n, d)
    er, es, ec = c_oracle.dense_topk(qq[:2], x, no.DT_F32, 10)
    assert rows[:2].tolist() == er.tolist() and np.array_equal(scores[:2], es)
    # ... and 64 queries through the batched (tensor-core) path against the pooled oracle: an fp32 GEMM proposes 256
    # rows per query, oracle.c re-scores them in the canonical fp64 order (the pool's depth is asserted inside)
    import bench
    q64 = synth.unit_queries(64, d, 2012)
    rows64, scores64, counts64 = c.topk(q64, 10)
    ei, es64 = bench.oracle_topk_full(x, q64, 10, "f32")
    assert (counts64 == 10).all() and rows64.tolist() == ei.tolist() and np.array_equal(scores64, es64)
    c.close()


# ------------------------------------------- tensor-core (tcgen05) path ----
@pytest.mark.parametrize("dtype,n,d,B", [("bf16", 50003, 1024, 300), ("f32", 30000, 1024, 129), ("f16", 8000, 512, 64),
                                         ("bf16", 700, 64, 9), ("f32", 256, 128, 128), ("bf16", 40000, 2048, 33)])
def test_dense_batched_tensor_core_path_vs_oracle(dtype, n, d, B):
    from b200rag import DeviceCorpus, _lib
    x = helpers.synth_unit(n, d, seed=n + 7)
    q = helpers.synth_unit(B, d, seed=n + 8)
    q[1] = x[n // 3]
    x[n // 5] = x[n // 7]                                  # an exact duplicate pair
    c = DeviceCorpus(d, dtype)
    c.append(x)
    before = _lib.counters()["fallbacks"]
    for k in (10, 1, 50, 100):
        check_topk(c, q, min(k, n), DT[dtype])
    allow = np.random.default_rng(3).random(n) < 0.25
    check_topk(c, q, 10, DT[dtype], allow)
    # appended rows invalidate the bf16 shadow
    c.append(helpers.synth_unit(100, d, seed=99))
    check_topk(c, q[:16], 10, DT[dtype])
    assert _lib.counters()["fallbacks"] - before <= 2      # random data: the margin check passes
    c.close()


@pytest.mark.parametrize("dtype,n,d,B", [("bf16", 50003, 1024, 512), ("f32", 40000, 512, 256), ("f16", 38000, 1024, 300),
                                         ("f16", 38100, 256, 256)])
def test_pair_mode_cta_group2_vs_oracle_and_single_cta(dtype, n, d, B):
    """Main pass as CTA pairs (tcgen05 cta_group::2): needs >= 148 row tiles and an even number of 128-query
    blocks (B = 300 -> 3 blocks: stays single-CTA).  Same results as the oracle and as the single-CTA kernel."""
    from b200rag import DeviceCorpus, _lib
    x = helpers.synth_unit(n, d, seed=n + 17)
    q = helpers.synth_unit(B, d, seed=n + 18)
    q[3] = x[n // 2]
    x[n - 1] = x[5]                                        # duplicate pair across the first and the last tile
    dup = np.random.default_rng(5).choice(n, size=700, replace=False)
    x[dup] = q[7]                                          # 700 exact ties: query 7 overflows into the fallback pass
    c = DeviceCorpus(d, dtype)
    c.append(x)
    allow = np.random.default_rng(4).random(n) < 0.5
    try:
        f0 = _lib.counters()["fallbacks"]
        r2, s2, _ = check_topk(c, q, 10, DT[dtype])
        assert r2[7].tolist() == sorted(dup.tolist())[:10] and _lib.counters()["fallbacks"] == f0 + 1
        ra2, sa2, _ = check_topk(c, q, 37, DT[dtype], allow)
        _lib.set_option("pair_mode", 0)
        r1, s1, _ = c.topk(q, 10)
        assert r1.tolist() == r2.tolist() and np.array_equal(s1, s2)
    finally:
        _lib.set_option("pair_mode", 1)
    c.close()


def test_tensor_core_path_exact_ties_take_the_fallback():
    from b200rag import DeviceCorpus, _lib
    n, d, B = 20000, 256, 40
    x = helpers.synth_unit(n, d, seed=21)
    q = helpers.synth_unit(B, d, seed=22)
    dup = np.random.default_rng(4).choice(n, size=900, replace=False)
    x[dup] = q[3]
    c = DeviceCorpus(d, "bf16")
    c.append(x)
    before = _lib.counters()["fallbacks"]
    rows, scores, counts = check_topk(c, q, 10, no.DT_BF16)
    assert rows[3].tolist() == sorted(dup.tolist())[:10]
    assert _lib.counters()["fallbacks"] == before + 1
    c.close()


def test_batched_front_end_equals_per_question_path(e2e_data, golden_dir):
    """retrieve_candidates_batch (one dense call + one BM25 call per filter + one device RRF call for many
    questions) returns what the reference returned question by question."""
    from b200rag import DeviceCollection, DeviceChunkBM25Index, DeviceSummaryBM25Index, HybridRetriever
    gold, emb, table = e2e_data
    col = DeviceCollection(dim=emb.shape[1], dtype="f32")
    helpers.fill(col, gold["chunks"], emb)
    bm = DeviceChunkBM25Index()
    bm.build_from_collection(col)
    sm = DeviceSummaryBM25Index()
    sm.build(os.path.join(golden_dir, "e2e_summaries.json"))
    groups = {}
    for run in gold["runs"]:
        key = (str(run["config"]), str(run["where"]))
        groups.setdefault(key, []).append(run)
    assert any(len(v) > 1 for v in groups.values())
    for runs in groups.values():
        cfg, where = runs[0]["config"], runs[0]["where"]
        r = HybridRetriever(collection=col, embedding_provider=helpers.FixedEmbeddingProvider(table),
                            summary_bm25_index=sm if cfg["prefilter"] else None, chunk_bm25_index=bm,
                            query_expander=helpers.FixedQueryExpander(gold["expansions"]) if cfg["expander"] else None,
                            summary_prefilter_k=8, enable_hybrid=cfg["hybrid"], enable_summary_prefilter=cfg["prefilter"],
                            acronym_expander=helpers.acronym_expander_for_golden(gold))
        got = r.retrieve_candidates_batch([run["query"] for run in runs], n_candidates=40, where_filter=where)
        for run, chunks in zip(runs, got):
            assert [helpers.chunk_dump(c) for c in chunks] == run["candidates"], (cfg, run["query"])


def test_scan_kernel_multi_query_variants_forced():
    """B = 2..7 through the CUDA-core scan kernel (NQ = 2 and 4 templates, several launches) by raising the
    tensor-core threshold; the default dispatch sends B >= 2 to the contraction kernel."""
    from b200rag import DeviceCorpus, _lib
    _lib.set_option("tc_min_batch", 1 << 20)
    _lib.set_option("tc_b1_shadow", 0)
    try:
        for dtype, n, d in [("bf16", 9000, 1024), ("f32", 5000, 512), ("f16", 3000, 256)]:
            x = helpers.synth_unit(n, d, seed=n)
            q = helpers.synth_unit(7, d, seed=n + 1)
            c = DeviceCorpus(d, dtype)
            c.append(x)
            for B, k in [(2, 10), (3, 50), (4, 100), (7, 10)]:
                check_topk(c, q[:B], k, DT[dtype])
            c.close()
    finally:
        _lib.set_option("tc_min_batch", 2)
        _lib.set_option("tc_b1_shadow", 1)


def test_single_query_on_fp32_corpus_uses_the_bf16_shadow_and_stays_exact():
    from b200rag import DeviceCorpus, _lib
    n, d = 300_000, 1024                      # >= 262144 rows: batch-1 goes through the contraction kernel
    c = DeviceCorpus(d, "f32")
    c.fill_synthetic(seed=5, nrows=n)
    q = helpers.synth_unit(3, d, seed=6)
    q[1] = c.download(12345, 1)[0]
    launches = _lib.counters()["launches"]
    for i in range(3):
        check_topk(c, q[i:i + 1], 10, no.DT_F32)
    check_topk(c, q[:1], 50, no.DT_F32)
    c.close()


# ------------------------------------------------- BASELINE full-size shapes ----
def _slice_exactness(c, q, rows, scores, lo, hi, dt):
    """Exactness restricted to the row slice [lo, hi): the rows of the slice whose oracle score reaches our k-th
    score must be exactly our returned rows that fall in the slice, with identical fp64 scores."""
    x = c.download(lo, hi - lo)
    raw = raw_rows(x, dt)
    for b in range(len(q)):
        osc = c_oracle.dense_scores(q[b], raw, dt)
        kth = scores[b, -1]
        kth_row = rows[b, -1]
        want = {int(lo + i): float(osc[i]) for i in np.nonzero(osc >= kth)[0]
                if osc[i] > kth or lo + i <= kth_row}
        got = {int(r): float(s) for r, s in zip(rows[b], scores[b]) if lo <= r < hi}
        assert got == want, (b, lo, hi)


def test_config5_shard_full_size():
    """BASELINE config 5, one of the 8 shards: 12.5M x 1024 bf16 (25.6 GB), top-10, batch 1 (scan kernel) and a
    batch through the tcgen05 kernel; both paths must agree and be exact on sampled row slices."""
    from b200rag import DeviceCorpus, synth
    n, d, k = 12_500_000, 1024, 10
    c = DeviceCorpus(d, "bf16", capacity=n)
    c.fill_synthetic(seed=1005, nrows=n)
    q = synth.unit_queries(6, d, 2005)
    planted = [7_654_321, n - 1]
    q[4:] = np.concatenate([c.download(r, 1) for r in planted])
    r1 = [c.topk(q[i:i + 1], k) for i in range(6)]                 # batch-1: CUDA-core scan
    rows = np.concatenate([r[0] for r in r1]); scores = np.concatenate([r[1] for r in r1])
    rows_b, scores_b, counts_b = c.topk(q, k)                      # batch-6: tcgen05 contraction
    assert np.array_equal(rows, rows_b) and np.array_equal(scores, scores_b)
    assert (counts_b == k).all() and (np.diff(scores, axis=1) <= 0).all()
    for j, r in enumerate(planted):
        assert rows[4 + j, 0] == r
    for b in range(6):                                             # slices that contain returned rows + a random one
        lo = int(rows[b, 0]) // 200_000 * 200_000
        _slice_exactness(c, q[b:b + 1], rows[b:b + 1], scores[b:b + 1], lo, min(n, lo + 200_000), no.DT_BF16)
    _slice_exactness(c, q[:3], rows[:3], scores[:3], 3_000_000, 3_300_000, no.DT_BF16)
    c.close()


def test_config3_shard_shape_batch4096_top100():
    """BASELINE config 3 shape per GPU at G=8: 1.25M x 1024 bf16, 4096-query batch, top-100."""
    from b200rag import DeviceCorpus, synth, _lib
    n, d, k, B = 1_250_000, 1024, 100, 4096
    c = DeviceCorpus(d, "bf16", capacity=n)
    c.fill_synthetic(seed=1003, nrows=n)
    q = synth.unit_queries(B, d, 2003)
    f0 = _lib.counters()["fallbacks"]
    rows, scores, counts = c.topk(q, k)
    assert (counts == k).all() and (np.diff(scores, axis=1) <= 0).all()
    assert _lib.counters()["fallbacks"] - f0 <= 1
    stored = c.download()
    pick = [0, 1, 2047, 4095]
    er, es, ec = c_oracle.dense_topk(q[pick], raw_rows(stored, no.DT_BF16), no.DT_BF16, k)
    assert rows[pick].tolist() == er.tolist() and np.array_equal(scores[pick], es)
    # 64 more queries against the pooled oracle (fp32 GEMM pool of 512 rows, canonical fp64 re-score by oracle.c)
    import bench
    pick64 = list(range(5, 4096, 64))
    ei, es64 = bench.oracle_topk_full(stored, q[pick64], k, "bf16", pool=512)
    assert rows[pick64].tolist() == ei.tolist() and np.array_equal(scores[pick64], es64)
    c.close()


def test_batch_larger_than_one_launch_is_sliced():
    from b200rag import DeviceCorpus
    n, d, B = 3000, 128, 4100
    x = helpers.synth_unit(n, d, seed=1)
    q = helpers.synth_unit(B, d, seed=2)
    c = DeviceCorpus(d, "bf16")
    c.append(x)
    rows, scores, counts = c.topk(q, 5)
    pick = [0, 4095, 4096, 4099]
    er, es, ec = c_oracle.dense_topk(q[pick], raw_rows(c.download(), no.DT_BF16), no.DT_BF16, 5)
    assert rows[pick].tolist() == er.tolist() and np.array_equal(scores[pick], es)
    c.close()


def test_bm25_mass_ties_and_many_ranges():
    """> 1024 documents tie for the best BM25 score (identical documents): ties must resolve to the lowest rows;
    also exercises a corpus with many 4096-row ranges."""
    from b200rag.bm25 import DeviceBM25, Postings
    n_docs = 30000
    g = np.random.default_rng(8)
    docs = [g.integers(10, 400, size=g.integers(5, 30)) for _ in range(n_docs)]
    same = np.array([1, 2, 3, 3, 7], dtype=np.int64)
    dup = np.sort(g.choice(n_docs, size=2500, replace=False))
    for r in dup:
        docs[r] = same.copy()
    # relabel to first-seen order
    flat = np.concatenate(docs)
    uniq, first = np.unique(flat, return_index=True)
    remap = np.full(int(flat.max()) + 1, -1, dtype=np.int64)
    remap[uniq[np.argsort(first)]] = np.arange(len(uniq))
    docs = [remap[d] for d in docs]
    p = Postings.from_term_ids(docs, n_terms=len(uniq))
    o = no.CsrBM25(docs)
    ix = DeviceBM25(p)
    q = remap[np.array([3, 7, 1])].astype(np.int32)
    for k in (10, 50, 224):
        rows, scores, counts = ix.search_ids([q, q[:1]], k)
        for i, qt in enumerate([q, q[:1]]):
            er, es = o.search(qt.tolist(), k)
            assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es)
    assert rows[0, :10].tolist() == dup[:10].tolist()
    ix.close()


@pytest.mark.parametrize("tile", [0, 4])
@pytest.mark.parametrize("dense_div", [0, 2, 8, 64, 1 << 20])
def test_bm25_filter_index_classes_vs_oracle(dense_div, tile):
    """The integer filter index holds a term three ways (untabled short list scanned whole, range-tabled run of the
    packed stream, dense 16-bit column): every mix must select exactly the oracle's rows, on the fast path.
    tile 0: the automatic plan (quarter-block tiles for this small corpus); tile 4: block tiles, i.e. the kernel that
    stages the runs through shared memory with TMA bulk copies (bm25_filter_tma_kernel)."""
    from b200rag import _lib
    from b200rag.bm25 import DeviceBM25, Postings
    n_docs = 21000                                      # 6 ranges, the last one ragged
    docs, n_terms = helpers.zipf_docs(n_docs, 3000, seed=77)
    g = np.random.default_rng(5)
    # a term present in every row, one clustered in a few rows of one range (a long run of a non-dense term),
    # and documents that are exact duplicates (ties)
    docs = [np.concatenate([d, [n_terms]]) for d in docs]
    for r in range(9000, 9700):
        docs[r] = np.concatenate([docs[r], [n_terms + 1] * int(g.integers(1, 4))])
    for r in (100, 4095, 4096, 20999):
        docs[r] = docs[7].copy()
    n_terms += 2
    p = Postings.from_term_ids(docs, n_terms=n_terms)
    o = no.CsrBM25(docs)
    _lib.set_option("bm25_dense_div", dense_div)
    try:
        ix = DeviceBM25(p)
    finally:
        _lib.set_option("bm25_dense_div", 8)
    df = np.diff(p.term_ptr)
    by_df = np.argsort(-df)
    allow = g.random(n_docs) < 0.3
    bitmap = np.packbits(allow, bitorder="little")
    queries = []
    for i in range(24):
        qt = np.concatenate([by_df[g.integers(0, 12, size=3)], by_df[g.integers(12, 300, size=4)],
                             g.integers(0, n_terms, size=g.integers(1, 6))]).astype(np.int32)
        if i % 3 == 0:
            qt = np.concatenate([qt, [n_terms - 2, -1, qt[0]]]).astype(np.int32)
        queries.append(qt)
    queries.append(np.array(list(docs[7][:6]) * 2, np.int32))               # the duplicated documents tie at the top
    if tile == 0:       # (a batch with such a query takes the direct-load kernel: keep the tile-4 batch on the TMA kernel)
        queries.append(by_df[:150].astype(np.int32))                        # more tokens than one pass of the kernel holds
    _lib.set_option("bm25_tile", tile)
    try:
        _check_bm25_classes(_lib, ix, o, queries, docs, n_terms, allow, bitmap, tile)
    finally:
        _lib.set_option("bm25_tile", 0)
    ix.close()


def _check_bm25_classes(_lib, ix, o, queries, docs, n_terms, allow, bitmap, tile):
    before = _lib.counters()["fallbacks"]
    rows, scores, counts = ix.search_ids(queries, 50)
    if tile == 0:                                       # (2 block tiles hold too few heads for some of these)
        assert _lib.counters()["fallbacks"] == before   # nothing of this was redone on the robust path
    # the clustered term: hundreds of rows of one tile tie exactly / hold the whole top-k -> flagged and redone on the
    # exact range path, still exact
    queries.append(np.array([n_terms - 1], np.int32))
    queries.append(np.concatenate([queries[0], [n_terms - 1]]).astype(np.int32))
    for k in (10, 50):
        rows, scores, counts = ix.search_ids(queries, k)
        rows1, scores1, counts1 = ix.search_ids(queries[:1], k)             # single-query call
        rows_f, scores_f, counts_f = ix.search_ids(queries, k, bitmap)
        for i, qt in enumerate(queries):
            er, es = o.search(qt.tolist(), k)
            assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es), (k, i)
            er, es = o.search(qt.tolist(), k, allow)
            assert rows_f[i, :counts_f[i]].tolist() == er.tolist() and np.array_equal(scores_f[i, :counts_f[i]], es), (k, i)
        assert rows1[0].tolist() == rows[0].tolist() and np.array_equal(scores1[0], scores[0])


@pytest.mark.parametrize("tile,tma,acc16", [(1, 1, 1), (2, 1, 1), (4, 1, 1), (4, 1, 0), (4, 0, 0)])
def test_bm25_filter_tile_sizes_vs_oracle(tile, tma, acc16):
    """the filter kernel's three tile sizes (4096 / 8192 / 16384 rows per CTA; picked by launch size in production,
    forced here) over column, run and scanned tokens, with and without a row filter, all equal to the oracle;
    block tiles through both kernels (tma 1: runs staged through shared memory by a producer warp, with 16-bit
    accumulators in a coarser unit — 4 CTAs per SM — or 32-bit ones; 0: direct loads)"""
    from b200rag import _lib
    from b200rag.bm25 import DeviceBM25, Postings
    n_docs = 40000
    docs, n_terms = helpers.zipf_docs(n_docs, 2000, seed=3, lo=10, hi=40)
    p = Postings.from_term_ids(docs, n_terms=n_terms)
    o = no.CsrBM25(docs)
    ix = DeviceBM25(p)
    g = np.random.default_rng(1)
    by_df = np.argsort(-np.diff(p.term_ptr))
    qs = [np.concatenate([by_df[g.integers(0, 10, size=3)], by_df[g.integers(10, 200, size=4)],
                          g.integers(0, n_terms, size=3)]).astype(np.int32) for _ in range(12)]
    allow = g.random(n_docs) < 0.4
    bm = np.packbits(allow, bitorder="little")
    _lib.set_option("bm25_tile", tile)
    _lib.set_option("bm25_tma", tma)
    _lib.set_option("bm25_acc16", acc16)
    try:
        for k in (10, 50):
            for mask, bits in ((None, None), (allow, bm)):
                rows, scores, counts = ix.search_ids(qs, k, bits)
                for i, qt in enumerate(qs):
                    er, es = o.search(qt.tolist(), k, mask)
                    assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es), (k, i)
    finally:
        _lib.set_option("bm25_tile", 0)
        _lib.set_option("bm25_tma", 1)
        _lib.set_option("bm25_acc16", 1)
    ix.close()


def test_bm25_16bit_accumulators_at_their_limit():
    """the batched block-tile kernel keeps 16-bit accumulators in a unit chosen from the launch's longest query, so that
    a query's tokens cannot sum to 2^16: repeat the term with the LARGEST idf*impact product 15 / 16 / 31 / 32 / 63 /
    64 / 100 times (duplicates count again, bm25_index.py:265 -> rank_bm25 get_scores) next to frequent (column) terms —
    every unit the kernel can pick and the hand-over to the 32-bit kernel, all bit-equal to the oracle"""
    from b200rag import _lib
    from b200rag.bm25 import DeviceBM25, Postings
    n_docs = 40000
    docs, n_terms = helpers.zipf_docs(n_docs, 2000, seed=13, lo=10, hi=40)
    # a rare term in short documents: large idf and large impact -> the largest products of the index
    rare = n_terms
    for r in (5, 16383, 16384, 20001, 39999):
        docs[r] = np.concatenate([docs[r][:3], [rare, rare, rare]])
    n_terms += 1
    p = Postings.from_term_ids(docs, n_terms=n_terms)
    o = no.CsrBM25(docs)
    ix = DeviceBM25(p)
    by_df = np.argsort(-np.diff(p.term_ptr))
    _lib.set_option("bm25_tile", 4)
    try:
        for reps in (15, 16, 31, 32, 63, 64, 100):
            qs = [np.array([rare] * reps, np.int32),
                  np.concatenate([by_df[:3], [rare] * (reps - 3)]).astype(np.int32),
                  np.concatenate([[rare] * (reps - 4), by_df[1:3], by_df[40:42]]).astype(np.int32)]
            rows, scores, counts = ix.search_ids(qs, 10)
            for i, qt in enumerate(qs):
                er, es = o.search(qt.tolist(), 10)
                assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es), (reps, i)
    finally:
        _lib.set_option("bm25_tile", 0)
    ix.close()


def test_clustered_corpus_through_the_tensor_core_path():
    """Rows sorted by cluster (adjacent rows are near-duplicates, as chunks of one document are): every query's
    neighbours sit in one or two row tiles, and the sample may miss the query's cluster entirely."""
    from b200rag import DeviceCorpus, _lib
    n_clusters, per, d = 600, 250, 256
    g = np.random.default_rng(3)
    centers = helpers.synth_unit(n_clusters, d, seed=4)
    x = np.repeat(centers, per, axis=0) + 0.08 * g.standard_normal((n_clusters * per, d)).astype(np.float32)
    x = no.l2_normalize_rows(x)
    q = no.l2_normalize_rows(centers[g.choice(n_clusters, size=40, replace=False)] +
                             0.05 * g.standard_normal((40, d)).astype(np.float32))
    for dtype in ("bf16", "f32"):
        c = DeviceCorpus(d, dtype)
        c.append(x)
        f0 = _lib.counters()["fallbacks"]
        check_topk(c, q, 10, DT[dtype])
        check_topk(c, q, 100, DT[dtype])
        check_topk(c, q[:1], 10, DT[dtype])
        assert _lib.counters()["fallbacks"] - f0 <= 1, "clustered rows should not push queries into the fallback pass"
        c.close()


def test_more_edges_k224_dim2048_zero_vectors_pinned_buffers():
    from b200rag import DeviceCorpus, DeviceCollection, pinned_empty, _lib
    # largest k on both paths, dim 2048
    n, d = 5000, 2048
    x = helpers.synth_unit(n, d, seed=31)
    q = helpers.synth_unit(6, d, seed=32)
    c = DeviceCorpus(d, "bf16")
    c.append(x)
    check_topk(c, q[:1], 224, no.DT_BF16)          # scan kernel, kp = 256
    check_topk(c, q, 224, no.DT_BF16)              # contraction kernel, kp = 512
    # caller-provided page-locked buffers are DMA'd directly and give the same answer
    qp = pinned_empty((6, d), np.float32)
    qp[:] = q
    out = (pinned_empty((6, 10), np.int32), pinned_empty((6, 10), np.float64), pinned_empty((6,), np.int32))
    r1, s1, c1 = c.topk(qp, 10, out=out)
    r2, s2, c2 = c.topk(q, 10)
    assert np.array_equal(r1, r2) and np.array_equal(s1, s2) and np.array_equal(c1, c2)
    c.close()
    # all-zero embedding rows / queries: cosine space normalises with a floor, scores are exactly 0
    col = DeviceCollection(dim=64, dtype="f32")
    emb = helpers.synth_unit(10, 64, seed=1)
    emb[3] = 0.0
    col.add(ids=[f"i{j}" for j in range(10)], documents=["t"] * 10, embeddings=emb, metadatas=[{"a": j} for j in range(10)])
    ora = no.ExactCollection(dim=64)
    ora.add(ids=[f"i{j}" for j in range(10)], documents=["t"] * 10, embeddings=emb, metadatas=[{"a": j} for j in range(10)])
    for qv in (emb[1].tolist(), [0.0] * 64):
        assert col.query(query_embeddings=[qv], n_results=10) == ora.query(query_embeddings=[qv], n_results=10)
    # update with a new embedding, get by ids / where
    col.update(ids=["i2"], embeddings=[emb[7].tolist()], metadatas=[{"a": 99}])
    ora.update(ids=["i2"], embeddings=[emb[7].tolist()], metadatas=[{"a": 99}])
    assert col.query(query_embeddings=[emb[7].tolist()], n_results=3) == ora.query(query_embeddings=[emb[7].tolist()], n_results=3)
    assert col.get(ids=["i5", "i2", "nope"])["ids"] == ora.get(ids=["i5", "i2", "nope"])["ids"]
    assert col.get(where={"a": {"$gte": 8}})["ids"] == ora.get(where={"a": {"$gte": 8}})["ids"]
    with pytest.raises(_lib.B200RagError):
        _lib.set_option("no_such_option", 1)


def test_opt_in_dense_doc_filter_prefilter(e2e_data):
    """N4: restricting the dense search to a set of documents BEFORE the top-k equals the exact search over only
    those rows (and differs from the reference's post-filter, which is why it is opt-in)."""
    from b200rag import DeviceCollection
    gold, emb, table = e2e_data
    col = DeviceCollection(dim=emb.shape[1], dtype="f32")
    helpers.fill(col, gold["chunks"], emb)
    docs = sorted({c["metadata"]["document_path"] for c in gold["chunks"]})[::4]
    q = np.array(list(table.values())[:3], dtype=np.float32)
    rows, scores, counts = col.query_rows(q, 25, where={"source": "CNIL"}, doc_filter=docs)
    allow = np.array([(c["metadata"]["document_path"] in set(docs)) and c["metadata"]["source"] == "CNIL"
                      for c in gold["chunks"]])
    stored = col.corpus.download()
    er, es, ec = c_oracle.dense_topk(no.l2_normalize_rows(q), stored, no.DT_F32, 25, np.packbits(allow, bitorder="little"))
    for b in range(3):
        assert rows[b, :counts[b]].tolist() == er[b, :ec[b]].tolist() and np.array_equal(scores[b, :counts[b]], es[b, :ec[b]])


def test_sharded_index_single_rank_device_path():
    """ShardedDenseIndex on one rank (no process group): device-resident local top-k + packed merge == corpus.topk"""
    from b200rag.sharded import ShardedDenseIndex
    n, d = 40000, 256
    idx = ShardedDenseIndex(d, n, dtype="bf16")
    idx.fill_synthetic(seed=3)
    q = helpers.synth_unit(9, d, seed=4)
    for B, k in [(1, 10), (9, 50)]:
        ids, scores, counts = idx.topk(q[:B], k)
        r, s, c = idx.corpus.topk(q[:B], k)
        assert ids.tolist() == r.astype(np.int64).tolist() and np.array_equal(scores, s) and counts.tolist() == c.tolist()


def test_concurrent_callers_share_one_collection(e2e_data):
    """Streamlit script threads share one cached instance (app.py:42): concurrent query() / search() calls must be
    serialised correctly."""
    import threading
    from b200rag import DeviceCollection, DeviceChunkBM25Index
    gold, emb, table = e2e_data
    col = DeviceCollection(dim=emb.shape[1], dtype="bf16")
    helpers.fill(col, gold["chunks"], emb)
    bm = DeviceChunkBM25Index()
    bm.build_from_collection(col)
    qs = [list(map(float, v)) for v in list(table.values())[:12]]
    texts = [c["text"][:60] for c in gold["chunks"][:12]]
    want_d = [col.query(query_embeddings=[q], n_results=20) for q in qs]
    want_b = [[(r.doc_key, r.score) for r in bm.search(t, top_k=20)] for t in texts]
    errors = []

    def worker(tid):
        try:
            for rep in range(5):
                for i in range(tid, 12, 4):
                    assert col.query(query_embeddings=[qs[i]], n_results=20) == want_d[i]
                    assert [(r.doc_key, r.score) for r in bm.search(texts[i], top_k=20)] == want_b[i]
        except Exception as e:     # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:2]


# ------------------------------------------------ round 2: tombstones, device predicates, shards, BM25 v2 ----
def test_tombstones_and_device_where_predicate(e2e_data):
    """delete = tombstones on the device (row numbers stay), `where` = a compiled program evaluated on the device
    into a cached bitmap; both against the oracle collection, incl. re-adding after deletes and compaction"""
    from b200rag import DeviceCollection
    gold, emb, table = e2e_data
    qs = [list(map(float, v)) for v in list(table.values())[:4]]
    wheres = [None, {"source": "CNIL"}, {"source": {"$ne": "ENTREPRISE"}}, {"tag_rh": True},
              {"$or": [{"source": "CNIL"}, {"$and": [{"source": "ENTREPRISE"}, {"tag_rh": True}]}]},
              {"chunk_nature": {"$in": ["GUIDE", "SANCTION"]}}, {"source": "NOPE"}, {"nope": {"$ne": 1}},
              {"chunk_index": {"$gte": 2}}]                       # the last one takes the host evaluator
    for dtype in ("bf16", "f32"):
        dev = DeviceCollection(dim=emb.shape[1], dtype=dtype)
        ora = no.ExactCollection(dim=emb.shape[1], dtype=DT[dtype])
        half = len(gold["chunks"]) // 2
        helpers.fill(dev, gold["chunks"][:half], emb[:half])
        helpers.fill(ora, gold["chunks"][:half], emb[:half])
        victims = [c["id"] for c in gold["chunks"][3:half:4]]
        dev.delete(ids=victims); ora.delete(ids=victims)
        assert dev.count() == ora.count() and dev.corpus.count() == half        # rows did not move
        assert dev.corpus.live_count() == ora.count()
        helpers.fill(dev, gold["chunks"][half:], emb[half:])                      # append after deletes
        helpers.fill(ora, gold["chunks"][half:], emb[half:])
        dev.delete(where={"tag_rh": True}); ora.delete(where={"tag_rh": True})
        some = dev.get(limit=3, offset=5)["ids"]
        dev.update(ids=some, metadatas=[{"source": "CNIL", "document_path": "x"}] * 3)
        ora.update(ids=some, metadatas=[{"source": "CNIL", "document_path": "x"}] * 3)
        for w in wheres:
            for n_res in (50, 7):
                assert dev.query(query_embeddings=qs[:2], n_results=n_res, where=w) == \
                    ora.query(query_embeddings=qs[:2], n_results=n_res, where=w), (dtype, w)
        got = dev.query(query_embeddings=qs[:1], n_results=5)
        got["metadatas"][0][0]["mutated"] = True                                  # results are fresh objects
        assert "mutated" not in dev.query(query_embeddings=qs[:1], n_results=5)["metadatas"][0][0]
        assert dev.get(limit=40, offset=7) ["ids"] == ora.get(limit=40, offset=7)["ids"]
        assert np.array_equal(dev.get(include=["embeddings"])["embeddings"], ora.get(include=["embeddings"])["embeddings"])
        dev.compact()                                                             # physical compaction keeps results
        assert dev.corpus.count() == ora.count()
        for w in wheres[:6]:
            assert dev.query(query_embeddings=qs[2:4], n_results=30, where=w) == \
                ora.query(query_embeddings=qs[2:4], n_results=30, where=w), (dtype, w)


def test_delete_is_cheap_and_exact_on_a_large_corpus():
    """100 tombstones on a 2M-row corpus: no row moves, the deleted rows never come back, results stay exact"""
    import time
    from b200rag import DeviceCorpus
    n, d = 2_000_000, 256
    c = DeviceCorpus(d, "bf16", capacity=n)
    c.fill_synthetic(seed=3, nrows=n)
    q = helpers.synth_unit(5, d, seed=4)
    rows0, scores0, _ = c.topk(q, 10)
    victims = np.unique(np.concatenate([rows0[:, :3].ravel(), np.arange(1000, 1000 + 80)]))
    t0 = time.perf_counter()
    c.delete_rows(victims)
    dt = time.perf_counter() - t0
    assert dt < 0.05, f"delete of {len(victims)} rows took {dt * 1e3:.1f} ms"
    assert c.count() == n and c.live_count() == n - len(victims)
    allow = np.ones(n, bool)
    allow[victims] = False
    for B in (1, 5):
        rows, scores, counts = c.topk(q[:B], 10)
        assert not np.isin(rows, victims).any()
        er, es, ec = c_oracle.dense_topk(q[:B], raw_rows(c.download(), no.DT_BF16), no.DT_BF16, 10,
                                         np.packbits(allow, bitorder="little"))
        assert rows.tolist() == er.tolist() and np.array_equal(scores, es)
    c.close()


def _sharded_devices(n_shards):
    import torch
    n_gpu = torch.cuda.device_count()
    return [i % n_gpu for i in range(n_shards)]


@pytest.mark.parametrize("dtype,n_shards", [("bf16", 3), ("f32", 2)])
def test_sharded_corpus_in_one_process_vs_oracle(dtype, n_shards):
    """rag_corpus_create_sharded: rows block-cyclic over shard slots (all GPUs of the box; with one GPU the slots
    share it), queries on every shard, (score, GLOBAL id) written into the primary shard's gather buffers, merged
    there.  Same results as the oracle over the whole corpus, incl. filters, tombstones and exact ties."""
    from b200rag import DeviceCorpus, _lib
    n, d = 23_456, 256
    x = helpers.synth_unit(n, d, seed=77)
    q = helpers.synth_unit(140, d, seed=78)
    x[20_000] = x[1500]                                    # a tie across shards: lowest global row wins
    q[1] = x[1500]
    dup = np.random.default_rng(5).choice(n, size=700, replace=False)
    x[dup] = q[7]                                          # deep ties: the exact fallback pass, per shard
    c = DeviceCorpus(d, dtype, capacity=n, n_shards=n_shards, devices=_sharded_devices(n_shards))
    c.append(x[:10_000])
    c.append(x[10_000:])                                   # appended in two uploads, not block-aligned
    assert c.count() == n
    assert np.array_equal(c.download(), no.quantize(x, DT[dtype]))
    assert np.array_equal(c.download(1000, 3000), no.quantize(x[1000:4000], DT[dtype]))
    for B, k in [(1, 10), (3, 50), (140, 10), (9, 100)]:
        check_topk(c, q[:B], k, DT[dtype])
    assert check_topk(c, q[7:8], 10, DT[dtype])[0][0].tolist() == sorted(dup.tolist())[:10]
    allow = np.random.default_rng(6).random(n) < 0.3
    check_topk(c, q[:5], 10, DT[dtype], allow)
    victims = np.unique(np.concatenate([np.arange(1500, 1600), dup[:300]]))
    c.delete_rows(victims)
    alive = np.ones(n, bool)
    alive[victims] = False
    rows, scores, counts = c.topk(q[:9], 10)
    er, es, ec = c_oracle.dense_topk(q[:9], raw_rows(c.download(), DT[dtype]), DT[dtype], 10, np.packbits(alive, bitorder="little"))
    assert rows.tolist() == er.tolist() and np.array_equal(scores, es)
    rows, scores, counts = c.topk(q[:4], 10, np.packbits(allow, bitorder="little"))
    er, es, ec = c_oracle.dense_topk(q[:4], raw_rows(c.download(), DT[dtype]), DT[dtype], 10, np.packbits(allow & alive, bitorder="little"))
    assert rows.tolist() == er.tolist() and np.array_equal(scores, es)
    c.close()


def test_sharded_collection_and_bm25_behind_the_drop_in(e2e_data, golden_dir):
    """DeviceCollection(n_shards=2) + DeviceChunkBM25Index(n_shards=2) inside ONE process return what the
    reference's retriever returned (the same golden as the single-GPU path)"""
    from b200rag import DeviceCollection, DeviceChunkBM25Index, DeviceSummaryBM25Index, HybridRetriever
    gold, emb, table = e2e_data
    devs = _sharded_devices(2)
    col = DeviceCollection(dim=emb.shape[1], dtype="f32", n_shards=2, devices=devs)
    helpers.fill(col, gold["chunks"], emb)
    bm = DeviceChunkBM25Index(n_shards=2, devices=devs)
    bm.build_from_collection(col)
    sm = DeviceSummaryBM25Index()
    sm.build(os.path.join(golden_dir, "e2e_summaries.json"))
    for run in gold["runs"][::3]:
        cands, docs = helpers.run_e2e_case(HybridRetriever, col, bm, sm, gold, table, run)
        assert cands == run["candidates"], (run["config"], run["query"])
        assert docs == run["documents"], (run["config"], run["query"])


def test_sharded_bm25_vs_oracle():
    from b200rag.bm25 import DeviceBM25, Postings
    n_docs, vocab = 40_000, 9000
    docs, n_terms = helpers.zipf_docs(n_docs, vocab, seed=11)
    p = Postings.from_term_ids(docs, n_terms=n_terms)
    o = no.CsrBM25(docs)
    g = np.random.default_rng(2)
    queries = [g.integers(0, 500, size=g.integers(2, 12)).astype(np.int32) for _ in range(20)]
    allow = g.random(n_docs) < 0.4
    few = np.zeros(n_docs, bool)
    few[g.choice(n_docs, size=300, replace=False)] = True        # selective filter: the listed-rows kernel
    for n_shards in (1, 3):
        ix = DeviceBM25(p, n_shards=n_shards, devices=_sharded_devices(n_shards))
        for k in (50, 10):
            for mask in (None, allow, few):
                bm = np.packbits(mask, bitorder="little") if mask is not None else None
                rows, scores, counts = ix.search_ids(queries, k, bm)
                for i, qt in enumerate(queries):
                    er, es = o.search(qt.tolist(), k, mask)
                    assert rows[i, :counts[i]].tolist() == er.tolist(), (n_shards, k, i)
                    assert np.array_equal(scores[i, :counts[i]], es)
        assert np.array_equal(ix.scores(queries[0]), o.get_scores(queries[0].tolist()))
        ix.close()


def test_bm25_config4_full_size_vs_oracle():
    """BASELINE config 4 size: 1M documents, Zipf(1.07) over 200k terms, 16 queries (8-12 Zipf terms + 2
    mid-frequency terms) against oracle.c over the same CSR: top-50 ids and fp64 scores bit-equal, batched and
    single-query paths, with and without a row filter; the packed filter path must be the one that served it"""
    import bench
    from b200rag import _lib
    from b200rag.bm25 import DeviceBM25, Postings
    doc_ptr, tokens, n_terms = bench.zipf_tokens(1_000_000, 200_000, 1004)
    post = Postings.from_flat_tokens(doc_ptr, tokens, n_terms)
    ix = DeviceBM25(post)
    assert ix.bytes_per_posting() == 4
    g = np.random.default_rng(2004)
    p = np.arange(1, n_terms + 1, dtype=np.float64) ** (-1.07)
    p /= p.sum()
    queries = []
    for _ in range(16):
        qt = g.choice(n_terms, size=g.integers(8, 13), p=p)
        queries.append(np.concatenate([qt, g.integers(n_terms // 100, n_terms // 10, size=2)]).astype(np.int32))
    queries[3] = np.concatenate([queries[3], queries[3][:2], [-1]]).astype(np.int32)      # repeats + unknown token
    allow = g.random(post.n_docs) < 0.2
    f0 = _lib.counters()["fallbacks"]
    for mask in (None, allow):
        bm = np.packbits(mask, bitorder="little") if mask is not None else None
        rows, scores, counts = ix.search_ids(queries, 50, bm)
        for i, qt in enumerate(queries):
            want = c_oracle.bm25_scores(post.term_ptr, post.post_row, post.post_tf, post.doc_len, post.idf, post.avgdl,
                                        post.k1, post.b, qt)
            er, es = c_oracle.bm25_select(want, 50, bm)
            assert rows[i, :counts[i]].tolist() == er.tolist() and np.array_equal(scores[i, :counts[i]], es), i
            if i < 4:
                r1, s1, c1 = ix.search_ids([qt], 50, bm)
                assert r1[0, :c1[0]].tolist() == er.tolist() and np.array_equal(s1[0, :c1[0]], es)
    assert _lib.counters()["fallbacks"] - f0 <= 2          # i.i.d. rows: the range heads cover the top-50
    ix.close()


def test_stream_ordered_dense_call_with_device_driven_fallback():
    """rag_dense_topk_dev only queues work; queries whose margin check fails (700 exact ties) are redone by the
    device-driven fallback pass inside the same stream-ordered call; more such queries than it serves are marked -1"""
    import torch
    from b200rag import DeviceCorpus
    n, d, B, k = 30_000, 256, 64, 10
    x = helpers.synth_unit(n, d, seed=31)
    q = helpers.synth_unit(B, d, seed=32)
    g = np.random.default_rng(9)
    tied = [5, 17, 40]
    dups = {}
    for b in tied:
        dups[b] = g.choice(n, size=700, replace=False)
        x[dups[b]] = q[b]
    for dtype in ("bf16", "f32"):
        c = DeviceCorpus(d, dtype)
        c.append(x)
        qd = torch.from_numpy(q).cuda()
        o_r = torch.full((B, k), -7, dtype=torch.int32, device="cuda")
        o_s = torch.zeros((B, k), dtype=torch.float64, device="cuda")
        o_c = torch.zeros((B,), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for _ in range(2):                                   # twice: the completion counter must reset itself
            c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
        torch.cuda.synchronize()
        er, es, ec = c_oracle.dense_topk(q, raw_rows(c.download(), DT[dtype]), DT[dtype], k)
        assert o_c.cpu().tolist() == ec.tolist()
        assert o_r.cpu().numpy().tolist() == er.tolist() and np.array_equal(o_s.cpu().numpy(), es)
        # batch-1 (scan kernel) through the same entry point
        c.topk_dev(qd[5:6].data_ptr(), 1, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
        torch.cuda.synchronize()
        assert o_r[0].cpu().tolist() == er[5].tolist()
        c.close()
    # more deep-tie queries in one call than the device-driven pass serves: the surplus is flagged, not wrong
    x2 = helpers.synth_unit(n, d, seed=33)
    tied2 = list(range(0, 12, 2))
    for b in tied2:
        x2[g.choice(n, size=700, replace=False)] = q[b]
    c = DeviceCorpus(d, "bf16")
    c.append(x2)
    c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
    torch.cuda.synchronize()
    counts = o_c.cpu().numpy()
    er, es, ec = c_oracle.dense_topk(q, raw_rows(c.download(), no.DT_BF16), no.DT_BF16, k)
    from b200rag import _lib
    f0 = _lib.counters()["fallback_queries"]
    rows, scores, _ = c.topk(q, k)                           # the host-buffer call serves all of them ...
    assert rows.tolist() == er.tolist() and np.array_equal(scores, es)
    n_flagged = _lib.counters()["fallback_queries"] - f0     # ... and tells how many queries needed the exact pass
    assert n_flagged >= len(tied2)
    assert (counts == -1).sum() == n_flagged - 4 and (counts[[8, 10]] == -1).all()
    good = counts >= 0
    assert o_r.cpu().numpy()[good].tolist() == er[good].tolist()
    assert np.array_equal(o_s.cpu().numpy()[good], es[good])
    c.close()


def test_large_k_and_large_fusion_do_not_degrade(e2e_data):
    """n_candidates above RAG_MAX_K (Chroma has no such limit): dense and BM25 serve it in passes, the batched
    front-end falls back to per-question calls when the rankings exceed one RRF call"""
    from b200rag import DeviceCorpus, reciprocal_rank_fusion
    from b200rag.bm25 import DeviceBM25, Postings
    n, d = 3000, 128
    x = helpers.synth_unit(n, d, seed=1)
    q = helpers.synth_unit(2, d, seed=2)
    c = DeviceCorpus(d, "f32")
    c.append(x)
    rows, scores, counts = c.topk(q, 500)
    er, es, ec = c_oracle.dense_topk(q, x, no.DT_F32, 500)
    assert rows.tolist() == er.tolist() and np.array_equal(scores, es)
    c.close()
    docs, n_terms = helpers.zipf_docs(3000, 800, seed=4)
    ix = DeviceBM25(Postings.from_term_ids(docs, n_terms=n_terms))
    o = no.CsrBM25(docs)
    qt = np.array([1, 5, 9, 40], np.int32)
    r, s, cnt = ix.search_ids([qt], 400)
    er, es = o.search(qt.tolist(), 400)
    assert r[0, :cnt[0]].tolist() == er.tolist() and np.array_equal(s[0, :cnt[0]], es)
    ix.close()
    rankings = [[f"id{(7 * i + r) % 900}" for i in range(700)] for r in range(6)]      # 4200 entries
    got = reciprocal_rank_fusion(rankings, k=60, weights=[2, 3, 1, 0.75, 1, 0.75])
    ids, sc = no.rrf_fuse(rankings, [2, 3, 1, 0.75, 1, 0.75], 60)
    assert got == dict(zip(ids, sc)) or {k: got[k] for k in ids} == dict(zip(ids, sc))


def test_batch_front_end_after_collection_mutation(e2e_data, golden_dir):
    """the batched front-end's BM25-row -> collection-row map must follow deletes / adds"""
    from b200rag import DeviceCollection, DeviceChunkBM25Index, HybridRetriever
    gold, emb, table = e2e_data
    col = DeviceCollection(dim=emb.shape[1], dtype="f32")
    helpers.fill(col, gold["chunks"], emb)
    bm = DeviceChunkBM25Index()
    bm.build_from_collection(col)
    runs = [r for r in gold["runs"] if r["where"] is None and r["config"]["hybrid"] and not r["config"]["prefilter"]][:4]
    assert runs
    cfg = runs[0]["config"]
    r = HybridRetriever(collection=col, embedding_provider=helpers.FixedEmbeddingProvider(table), chunk_bm25_index=bm,
                        query_expander=helpers.FixedQueryExpander(gold["expansions"]) if cfg["expander"] else None,
                        enable_hybrid=True, enable_summary_prefilter=False,
                        acronym_expander=helpers.acronym_expander_for_golden(gold))
    qs = [x["query"] for x in runs]
    r.retrieve_candidates_batch(qs, n_candidates=40)             # builds the row map
    victims = [c["id"] for c in gold["chunks"][10:200:7]]
    col.delete(ids=victims)
    col.compact()                                                # rows shift
    for stale in (True, False):
        if not stale:
            bm.build_from_collection(col)                        # rebuilt index: fast batch path again
        got = r.retrieve_candidates_batch(qs, n_candidates=40)
        want = [r.retrieve_candidates(x, n_candidates=40) for x in qs]
        for a, b in zip(got, want):
            assert [helpers.chunk_dump(c) for c in a] == [helpers.chunk_dump(c) for c in b]
            if not stale:
                assert not set(c.chunk_id for c in a) & set(victims)


# ------------------------------------------------------------- rerank ----
def test_rerank_step_golden_from_reference_and_batched_vs_oracle():
    """DeviceRerankStep.rerank == the reference's CrossEncoderReranker.rerank (golden: its own code around a table
    scorer): same pairs handed to the scorer, same chunks, final scores bit-equal, same error on the empty result;
    rerank_batch == question by question; random batches of the select kernel == the oracle restatement."""
    from b200rag import DeviceRerankStep, RetrievedChunk, rerank_select
    from oracle import gen_golden, ref_harness
    cases = gen_golden.rerank_cases()
    gold = load_golden("rerank.json")
    tm = ref_harness.TagTopicMatcher()

    def build(c):
        chunks = [RetrievedChunk(chunk_id=d["chunk_id"], text=d["text"], document_path=d["document_path"], chunk_nature="GUIDE",
                                 chunk_index=d["metadata"]["chunk_index"], confidence="high", distance=d["distance"],
                                 metadata=d["metadata"]) for d in c["chunks"]]
        table = {}
        for d, s in zip(c["chunks"], c["model_scores"]):
            text = d["text"]
            if d["metadata"].get("heading", ""):
                text = f"{d['metadata']['heading']}\n{text}"
            table[(c["query"], text[:512 * 4])] = np.float32(s)
        return chunks, table

    for c, g in zip(cases, gold):
        chunks, table = build(c)
        scorer = ref_harness.TableScorer(table)
        step = DeviceRerankStep(scorer, min_score=c["min_score"])
        if "raises" in g:
            with pytest.raises(IndexError):
                step.rerank(c["query"], chunks, top_k=c["top_k"], topic_matcher=tm, question_topics=c["topics"])
            continue
        got = step.rerank(c["query"], chunks, top_k=c["top_k"], topic_matcher=tm, question_topics=c["topics"])
        assert (scorer.calls[0] if scorer.calls else []) == g["pairs"]
        assert [(r.chunk_id, float(r.rerank_score).hex(), r.original_rank) for r in got] == \
            [(r["chunk_id"], r["rerank_score"], r["original_rank"]) for r in g["result"]]
    # all questions with the same top_k / min_score in one call
    same = [(c, g) for c, g in zip(cases, gold) if c["top_k"] == 10 and c["min_score"] == 0.08 and "raises" not in g]
    assert len(same) >= 2
    table = {}
    lists = []
    for c, _ in same:
        chunks, t = build(c)
        table.update(t)
        lists.append(chunks)
    scorer = ref_harness.TableScorer(table)
    step = DeviceRerankStep(scorer, min_score=0.08)
    out = step.rerank_batch([c["query"] for c, _ in same], lists, top_k=10, topic_matcher=tm,
                            question_topics_list=[c["topics"] for c, _ in same])
    assert len(scorer.calls) == 1                       # ONE scorer call for the whole batch
    for res, (_, g) in zip(out, same):
        assert [(r.chunk_id, float(r.rerank_score).hex()) for r in res] == [(r["chunk_id"], r["rerank_score"]) for r in g["result"]]
    # the kernel against the oracle: ragged lengths, ties, boosts, thresholds
    rng = np.random.default_rng(9)
    Q, L = 200, 64
    sc = rng.choice(np.linspace(0, 1, 23), size=(Q, L)).astype(np.float32)
    bo = np.where(rng.random((Q, L)) < 0.2, rng.choice([0.15, 0.05, 0.1], size=(Q, L)), 0.0)
    lens = rng.integers(0, L + 1, size=Q).astype(np.int32)
    for top_k, ms in ((8, 0.08), (10, 0.6), (40, 0.3), (3, 2.0)):
        idx, final, counts = rerank_select(sc, bo, lens, top_k, ms)
        for q in range(Q):
            ei, ef = no.rerank_select(sc[q, :lens[q]], bo[q, :lens[q]], top_k, ms)
            assert idx[q, :counts[q]].tolist() == ei and np.array_equal(final[q, :counts[q]], np.array(ef)), (top_k, q)
            assert (idx[q, counts[q]:] == -1).all()


def test_collection_loaded_from_a_chroma_persist_directory(tmp_path):
    """DeviceCollection built by b200rag.chroma_store.load_collection from a Chroma persist directory answers like
    the collection the same rows were add()ed to (cosine space: the stored unit vectors and the raw log vectors
    normalise to the same rows up to the last bit of the normalisation, so ids are compared, distances to 1e-6)"""
    from b200rag import DeviceCollection
    from b200rag.chroma_store import load_collection
    from oracle import chroma_fixture as cf
    g = np.random.default_rng(12)
    n, dim = 3000, 128
    emb = helpers.synth_unit(n, dim, seed=21) * g.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    ids = [f"c{i}" for i in range(n)]
    docs = [f"document {i}" for i in range(n)]
    metas = [{"document_path": f"doc{i % 37}", "chunk_nature": ["GUIDE", "FAQ", "LOI"][i % 3], "chunk_index": i} for i in range(n)]
    cf.write_store(str(tmp_path), ids, docs, metas, emb, n_flushed=2000, deleted_labels=(777777,))
    col = load_collection(str(tmp_path), dtype="f32")
    ref = DeviceCollection(dim=dim, dtype="f32")
    ref.add(ids=ids, documents=docs, embeddings=emb, metadatas=metas)
    assert col.count() == n
    q = helpers.synth_unit(6, dim, seed=22)
    for where in (None, {"chunk_nature": "FAQ"}, {"document_path": {"$in": ["doc3", "doc5"]}}):
        a = col.query(query_embeddings=q.tolist(), n_results=10, where=where)
        b = ref.query(query_embeddings=q.tolist(), n_results=10, where=where)
        assert a["ids"] == b["ids"] and a["documents"] == b["documents"] and a["metadatas"] == b["metadatas"]
        assert np.allclose(np.array(a["distances"]), np.array(b["distances"]), atol=1e-6)
    got = col.get(ids=["c5", "c2500"])
    assert got["documents"] == ["document 5", "document 2500"]
