"""Oracle pinning (CPU): the restatements in oracle/ against the golden vectors
produced by the reference's own Python (oracle/gen_golden.py) and against each
other (numpy vs C)."""
import numpy as np
import pytest

from conftest import load_golden, unhex
from oracle import c_oracle
from oracle import numpy_oracle as no
from oracle import rank_bm25 as rb
from oracle import ref_harness
import helpers


def test_rrf_known_answer_from_reference():
    # G1: computed by reciprocal_rank_fusion (src/rag/retriever.py:66-90) itself
    ids, sc = no.rrf_fuse([["a", "b"], ["b", "c"]], weights=[2.0, 3.0])
    got = dict(zip(ids, sc))
    assert got == {"a": 0.03278688524590164, "b": 0.08143839238498149, "c": 0.04838709677419355}


def test_rrf_golden_numpy_and_c():
    for case in load_golden("rrf.json"):
        ids, sc = no.rrf_fuse(case["rankings"], case["weights"], case["k"])
        assert ids == case["order"]
        assert [s.hex() for s in sc] == [case["scores"][i] for i in ids]
        # C restatement over integer ids
        names = {}
        R = len(case["rankings"])
        L = max(1, max(len(r) for r in case["rankings"]))
        arr = np.full((R, L), -1, np.int32)
        for r, ranking in enumerate(case["rankings"]):
            for j, key in enumerate(ranking):
                arr[r, j] = names.setdefault(key, len(names))
        inv = {v: k for k, v in names.items()}
        w = case["weights"] or [1.0] * R
        ci, cs = c_oracle.rrf(arr, w, case["k"])
        assert [inv[i] for i in ci] == case["order"]
        assert [float(s).hex() for s in cs] == [case["scores"][i] for i in case["order"]]


def test_canonical_score_numpy_equals_c_and_fp64_matmul(dense_small):
    x, q, _ = dense_small
    for dt in (no.DT_F32, no.DT_BF16, no.DT_F16):
        xs = no.quantize(x, dt)
        raw = xs if dt == no.DT_F32 else (no.f32_to_bf16_bits(x) if dt == no.DT_BF16 else x.astype(np.float16).view(np.uint16))
        for qi in range(3):
            a = no.canonical_scores(q[qi], xs)
            b = c_oracle.dense_scores(q[qi], raw, dt)
            assert np.array_equal(a, b)
            ref = xs.astype(np.float64) @ q[qi].astype(np.float64)
            assert np.max(np.abs(a - ref)) < 1e-14


def test_dense_golden(dense_small):
    x, q, gold = dense_small
    n = gold["n"]
    allow = (np.arange(n) % 3 != 0)
    bitmap = np.packbits(allow, bitorder="little")
    dts = {"f32": no.DT_F32, "bf16": no.DT_BF16, "f16": no.DT_F16}
    for case in gold["cases"]:
        dt = dts[case["dtype"]]
        xs = no.quantize(x, dt)
        raw = xs if dt == no.DT_F32 else (no.f32_to_bf16_bits(x) if dt == no.DT_BF16 else x.astype(np.float16).view(np.uint16))
        k = case["k"]
        cr, cs, cc = c_oracle.dense_topk(q, raw, dt, k, bitmap if case["filtered"] else None)
        for qi in range(len(q)):
            if qi < 3:   # numpy restatement is slow-ish: spot check
                r, s = no.dense_topk(q[qi], xs, k, allow if case["filtered"] else None)
                assert [int(v) for v in r] == case["rows"][qi]
                assert [float(v).hex() for v in s] == case["scores"][qi]
            assert cr[qi, :cc[qi]].tolist() == case["rows"][qi]
            assert [float(v).hex() for v in cs[qi, :cc[qi]]] == case["scores"][qi]


def test_dense_ties_lowest_row(dense_small):
    x, q, gold = dense_small
    # rows 5, 13, 700, 701 are identical and q[1] equals them: ties must come out in row order
    r, s = no.dense_topk(q[1], x, 4)
    assert r.tolist() == [5, 13, 700, 701] and len(set(s.tolist())) == 1


def test_csr_bm25_equals_dict_bm25_bitwise():
    docs, vocab = helpers.zipf_docs(400, 300, seed=4, lo=5, hi=60)
    words = [[f"w{t}" for t in d] for d in docs]
    ref = rb.BM25Okapi(words)
    csr = no.CsrBM25(docs)
    assert csr.avgdl == ref.avgdl
    for t in range(vocab):
        assert csr.idf[t] == ref.idf[f"w{t}"]
    assert (csr.idf < 0).any() or True
    g = np.random.default_rng(0)
    for _ in range(10):
        qt = g.integers(0, vocab + 5, size=g.integers(1, 12)).tolist()
        a = ref.get_scores([f"w{t}" for t in qt])
        b = csr.get_scores(qt)
        qt_c = [t if t < vocab else -1 for t in qt]
        c = c_oracle.bm25_scores(csr.term_ptr, csr.post_row, csr.post_tf, csr.doc_len, csr.idf, csr.avgdl, 1.5, 0.75, qt_c)
        assert np.array_equal(a, b) and np.array_equal(a, c)
        r1, s1 = csr.search(qt, 50)
        r2, s2 = c_oracle.bm25_select(c, 50)
        assert r1.tolist() == r2.tolist() and np.array_equal(s1, s2)


def test_bm25_negative_idf_floor():
    # a term in > half of the documents gets idf < 0 -> replaced by 0.25 * average idf
    docs = [["commun", "rare%d" % i] for i in range(10)] + [["autre", "mot"]]
    m = rb.BM25Okapi(docs)
    raw = np.log(11 - 10 + 0.5) - np.log(10 + 0.5)
    assert raw < 0 and m.idf["commun"] == 0.25 * m.average_idf and m.idf["commun"] != raw
    # all-common corpus: every idf negative, the floor is negative, nothing scores > 0
    m2 = rb.BM25Okapi([["a", "b"]] * 6)
    assert (m2.get_scores(["a", "b"]) <= 0).all()


def _oracle_chunk_index(chunks):
    """ChunkBM25Index.build_from_collection semantics (bm25_index.py:190-239) with the restated
    tokenizer-free path: uses the reference tokens recorded in the golden via rank_bm25 restatement."""
    from b200rag.tokenizer import tokenize_french
    kept, toks = [], []
    for c in chunks:
        if not c["text"] or not c["text"].strip():
            continue
        t = tokenize_french(c["text"])
        if not t:
            continue
        kept.append(c)
        toks.append(t)
    return kept, toks


def test_bm25_golden_from_reference_index():
    gold = load_golden("bm25_small.json")
    from b200rag.tokenizer import tokenize_french
    kept, toks = _oracle_chunk_index(gold["chunks"])
    assert [c["id"] for c in kept] == gold["kept_ids"]
    model = rb.BM25Okapi(toks)
    assert float(model.avgdl).hex() == gold["avgdl"]
    assert {w: float(v).hex() for w, v in model.idf.items()} == gold["idf"]
    for case in gold["cases"]:
        qt = tokenize_french(case["query"])
        if not qt:
            assert case["results"] == []
            continue
        scores = model.get_scores(qt)
        scored = [(i, scores[i]) for i in range(len(scores)) if scores[i] > 0 and
                  (case["doc_filter"] is None or kept[i]["metadata"]["document_path"] in set(case["doc_filter"]))]
        scored.sort(key=lambda t: t[1], reverse=True)
        got = [{"doc_key": kept[i]["id"], "score": float(s).hex()} for i, s in scored[:case["top_k"]]]
        assert got == case["results"]
    assert gold["common_results"] == []


@pytest.mark.needs_reference
@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present (GPU box)")
def test_goldens_are_reproducible_from_the_reference():
    ref = ref_harness.load()
    f = ref["retriever"].reciprocal_rank_fusion
    for case in load_golden("rrf.json"):
        kw = {"k": case["k"]}
        if case["weights"] is not None:
            kw["weights"] = case["weights"]
        assert {i: float(s).hex() for i, s in f(case["rankings"], **kw).items()} == case["scores"]
    tok = ref["bm25_index"].tokenize_french
    for item in load_golden("tokenizer.json"):
        assert tok(item["text"]) == item["tokens"]


@pytest.mark.needs_reference
@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present (GPU box)")
def test_reference_retriever_runs_around_the_oracle(e2e_data):
    """The reference's unmodified RAGRetriever + ChunkBM25Index around ExactCollection reproduce the e2e golden."""
    gold, emb, table = e2e_data
    ref = ref_harness.load()
    col = no.ExactCollection(dim=emb.shape[1])
    helpers.fill(col, gold["chunks"], emb)
    bm = ref["bm25_index"].ChunkBM25Index()
    bm.build_from_collection(col)
    run = gold["runs"][1]
    r = ref["retriever"].RAGRetriever(collection=col, llm_provider=None,
                                      embedding_provider=ref_harness.FixedEmbeddingProvider(table),
                                      chunk_bm25_index=bm, n_documents=5, n_chunks_per_doc=3, summary_prefilter_k=8,
                                      enable_hybrid=run["config"]["hybrid"], enable_summary_prefilter=False)
    got = r.retrieve_candidates(run["query"], n_candidates=40, where_filter=run["where"])
    assert [helpers.chunk_dump(c) for c in got] == run["candidates"]


def test_restated_hnsw_index_recall_against_exact_search():
    """The approximate store the reference queries (chromadb -> HNSW), restated in oracle/hnsw.c: on a clustered
    corpus (what document chunks look like) it finds almost all of the exact top-10; its distances are exact."""
    g = np.random.default_rng(5)
    n_clusters, per, d = 80, 50, 64
    centers = no.l2_normalize_rows(g.standard_normal((n_clusters, d)).astype(np.float32))
    x = no.l2_normalize_rows(np.repeat(centers, per, axis=0) + 0.25 * g.standard_normal((n_clusters * per, d)).astype(np.float32))
    q = no.l2_normalize_rows(centers[:20] + 0.2 * g.standard_normal((20, d)).astype(np.float32))
    ix = c_oracle.HnswIndex(x, M=16, ef_construction=100)
    ids, dist = ix.query(q, 10, ef_search=100)
    er, es, ec = c_oracle.dense_topk(q, x, no.DT_F32, 10)
    hits = sum(len(set(ids[i].tolist()) & set(er[i].tolist())) for i in range(len(q)))
    assert hits / (10 * len(q)) >= 0.9
    for i in range(len(q)):
        assert (np.diff(dist[i]) >= 0).all()
        np.testing.assert_allclose(dist[i], 1.0 - (x[ids[i]] @ q[i]), atol=2e-6)
    ix.close()


def test_rerank_select_golden_from_reference():
    """oracle restatement of the rerank tail == the reference's CrossEncoderReranker.rerank around a table scorer"""
    from oracle import gen_golden, ref_harness
    cases = gen_golden.rerank_cases()
    gold = load_golden("rerank.json")
    assert len(cases) == len(gold)
    tm = ref_harness.TagTopicMatcher()
    for c, g in zip(cases, gold):
        boosts = [tm.topic_boost(c["topics"], d["metadata"].get("rgpd_topics", "")) if c["topics"] else 0.0
                  for d in c["chunks"]]
        idx, final = no.rerank_select(c["model_scores"], boosts, c["top_k"], c["min_score"])
        if "raises" in g:
            assert idx == []
            continue
        assert [c["chunks"][i]["chunk_id"] for i in idx] == [r["chunk_id"] for r in g["result"]]
        assert [float(f).hex() for f in final] == [r["rerank_score"] for r in g["result"]]
        assert idx == [r["original_rank"] for r in g["result"]]
