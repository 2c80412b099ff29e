"""CPU tests of the host side: the C-ABI library loads and exports every symbol
include/b200rag.h declares, the host-side mirror of the reference interface
behaves like the reference (golden vectors), and nothing computes without a GPU."""
import os
import re

import numpy as np
import pytest

import helpers
from conftest import ROOT, load_golden, unhex
from oracle import numpy_oracle as no


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200rag.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rag_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from b200rag import _lib
    L = _lib.load()                    # dlopen only: no GPU needed
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/b200rag.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert L.rag_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from b200rag import DeviceCollection, B200RagError, reciprocal_rank_fusion
    with pytest.raises(B200RagError):
        DeviceCollection(dim=64)
    with pytest.raises(B200RagError):
        reciprocal_rank_fusion([["a"], ["b"]])


def test_no_product_import_of_oracle():
    pkg = os.path.join(ROOT, "rag-dpo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_tokenizer_matches_reference_golden():
    from b200rag import tokenize_french
    for item in load_golden("tokenizer.json"):
        assert tokenize_french(item["text"]) == item["tokens"], item["text"]


def test_where_matches_oracle():
    from b200rag.where import match, bitmap_from_mask
    metas = [{"source": "CNIL", "n": 1}, {"source": "ENTREPRISE", "tag_rh": True}, {"source": "ENTREPRISE"},
             {}, None, {"source": "CNIL", "tag_rh": False, "chunk_nature": "GUIDE", "n": 3}]
    wheres = [None, {}, {"source": "CNIL"}, {"tag_rh": True}, {"source": {"$ne": "ENTREPRISE"}},
              {"chunk_nature": {"$in": ["GUIDE", "X"]}}, {"$and": [{"source": "ENTREPRISE"}, {"tag_rh": True}]},
              {"$or": [{"source": "CNIL"}, {"$and": [{"source": "ENTREPRISE"}, {"tag_rh": True}]}]},
              {"n": {"$gte": 2}}, {"source": {"$nin": ["CNIL"]}}, {"tag_rh": 1}]
    for w in wheres:
        assert [match(m, w) for m in metas] == [no.where_match(m or {}, w) for m in metas], w
    assert bitmap_from_mask([1, 0, 0, 0, 0, 0, 0, 0, 1]).tolist() == [1, 1]


def test_postings_match_oracle_and_reference_idf():
    from b200rag.bm25 import Postings
    from b200rag.tokenizer import tokenize_french
    gold = load_golden("bm25_small.json")
    toks = [tokenize_french(c["text"]) for c in gold["chunks"] if c["text"].strip() and tokenize_french(c["text"])]
    p = Postings.from_token_lists(toks)
    assert float(p.avgdl).hex() == gold["avgdl"]
    assert {w: float(p.idf[t]).hex() for w, t in p.vocab.items()} == gold["idf"]
    # CSR identical to the oracle's
    ids = [np.array([p.vocab[w] for w in t]) for t in toks]
    o = no.CsrBM25(ids)
    assert np.array_equal(o.term_ptr, p.term_ptr) and np.array_equal(o.post_row, p.post_row)
    assert np.array_equal(o.post_tf, p.post_tf) and np.array_equal(o.idf, p.idf)
    assert p.term_ids(["données", "zzzz"]).tolist() == [p.vocab["données"], -1]


def test_hybrid_retriever_host_logic_reproduces_reference_golden(e2e_data, golden_dir):
    """HybridRetriever (host mirror of RAGRetriever) around CPU checkers == what the reference's own
    RAGRetriever returned (tests/golden/e2e_retrieve.json)."""
    from b200rag import HybridRetriever, tokenize_french
    gold, emb, table = e2e_data
    col = no.ExactCollection(dim=emb.shape[1])
    helpers.fill(col, gold["chunks"], emb)
    bm = helpers.OracleChunkBM25Index(tokenize_french)
    bm.build_from_collection(col)
    sm = helpers.OracleSummaryBM25Index(tokenize_french)
    sm.build(os.path.join(golden_dir, "e2e_summaries.json"))
    for run in gold["runs"]:
        cands, docs = helpers.run_e2e_case(HybridRetriever, col, bm, sm, gold, table, run, fuse=helpers.oracle_fuse)
        assert cands == run["candidates"], (run["config"], run["query"])
        assert docs == run["documents"], (run["config"], run["query"])


def test_summary_index_golden_cpu_checker(golden_dir):
    from b200rag import tokenize_french
    gold = load_golden("summary_bm25.json")
    sm = helpers.OracleSummaryBM25Index(tokenize_french)
    sm.build(os.path.join(golden_dir, "summaries_input.json"))
    assert sm.doc_keys == gold["doc_keys"]
    for case in gold["cases"]:
        got = [{"doc_key": k, "score": float(s).hex()} for k, s in sm.search_pairs(case["query"], case["top_k"])]
        assert got == case["results"]


def test_shard_bounds_cover_all_rows():
    from b200rag.sharded import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_where_program_equals_host_evaluator():
    """the compiled predicate (what csrc/rowfilter.cu evaluates on the device) against match() on every `where`
    shape the reference emits, via the host interpreter of the program"""
    from b200rag.where import ColumnCodes, Unsupported, compile_where, match, run_program
    metas = [{"source": "CNIL", "n": 1, "document_path": "a"}, {"source": "ENTREPRISE", "tag_rh": True, "document_path": "b"},
             {"source": "ENTREPRISE", "document_path": "b"}, {}, None,
             {"source": "CNIL", "tag_rh": False, "chunk_nature": "GUIDE", "n": 3, "document_path": "c"},
             {"source": "CNIL", "tag_rh": 1, "chunk_nature": ["list"], "document_path": "a"}]
    cols = ColumnCodes()
    coded = cols.encode_batch(metas)
    def codes_of(row):
        return lambda col: int(coded[col][row]) if col in coded else -1
    wheres = [None, {}, {"source": "CNIL"}, {"tag_rh": True}, {"source": {"$ne": "ENTREPRISE"}}, {"tag_rh": 1},
              {"chunk_nature": {"$in": ["GUIDE", "X"]}}, {"$and": [{"source": "ENTREPRISE"}, {"tag_rh": True}]},
              {"$or": [{"source": "CNIL"}, {"$and": [{"source": "ENTREPRISE"}, {"tag_rh": True}]}]},
              {"source": {"$nin": ["CNIL"]}}, {"nope": "x"}, {"nope": {"$ne": "x"}}, {"source": "NOPE"},
              {"source": {"$in": []}}, {"source": "CNIL", "tag_rh": False},
              {"$or": [{"tag_rh": {"$ne": True}}, {"n": 3}]}, {"source": {"$eq": "CNIL", "$ne": "X"}}]
    for w in wheres:
        for docs in (None, {"a", "zzz"}, set()):
            prog = compile_where(w, cols, docs)
            for r, m in enumerate(metas):
                want = match(m, w) and (docs is None or (m or {}).get("document_path", "") in docs)
                got = True if prog is None else run_program(prog, codes_of(r))
                assert got == want, (w, docs, r)
    with pytest.raises(Unsupported):
        compile_where({"n": {"$gte": 2}}, cols)


def test_native_csr_builder_matches_oracle():
    """rag_csr_build (multi-threaded host code in the library, no GPU) == the oracle's numpy CSR"""
    from b200rag.bm25 import Postings
    for n_docs, vocab, threads in [(3000, 900, 0), (257, 50, 3), (1, 5, 1)]:
        docs, n_terms = helpers.zipf_docs(n_docs, vocab, seed=n_docs, lo=0 if n_docs > 1 else 3, hi=60)
        lens = np.array([len(d) for d in docs])
        doc_ptr = np.concatenate([[0], np.cumsum(lens)])
        flat = np.concatenate(docs).astype(np.int32) if lens.sum() else np.zeros(0, np.int32)
        p = Postings.from_flat_tokens(doc_ptr, flat, n_terms, threads=threads)
        o = no.CsrBM25(docs)
        v = len(o.term_ptr) - 1
        assert np.array_equal(p.term_ptr[:v + 1], o.term_ptr) and (p.term_ptr[v:] == p.term_ptr[v]).all()
        assert np.array_equal(p.post_row, o.post_row) and np.array_equal(p.post_tf, o.post_tf)
        assert np.array_equal(p.doc_len, o.doc_len) and np.array_equal(p.idf[:v], o.idf)
    with pytest.raises(ValueError):
        Postings.from_flat_tokens(np.array([0, 2]), np.array([0, 7], np.int32), 3)


def test_block_cyclic_shard_layout_roundtrip():
    """the block-cyclic layout of sharded corpora (RAG_SHARD_BLOCK rows per block), python twin of api_common.h"""
    BS = 1024
    for G in (1, 2, 3, 8):
        n = 5 * BS * G + 17
        rows = np.arange(n)
        shard = (rows // BS) % G if G > 1 else np.zeros(n, int)
        local = (rows // BS // G) * BS + rows % BS if G > 1 else rows
        back = ((local // BS) * G + shard) * BS + local % BS if G > 1 else local
        assert np.array_equal(back, rows)
        for s in range(G):
            ls = local[shard == s]
            assert np.array_equal(ls, np.arange(len(ls)))          # local rows are dense and in global order


def test_chroma_store_reader_roundtrip(tmp_path):
    """the loader of the reference's on-disk vector store (chroma.sqlite3 + HNSW segment; scripts/package_cnil_db.py
    ships exactly these files) against a fixture written in the same layout: ids / documents / typed metadata in
    insertion order, flushed vectors from data_level0.bin (deleted elements skipped), the unflushed tail from the
    write-ahead log (stale records ignored).  FORMAT PARITY UNPINNED (no chromadb wheel here): both sides restate it."""
    from oracle import chroma_fixture as cf
    from b200rag import chroma_store as cs
    g = np.random.default_rng(4)
    n, dim = 70, 64
    emb = g.standard_normal((n, dim)).astype(np.float32)
    ids = [f"chunk_{i}" for i in range(n)]
    docs = [f"texte {i} é" if i % 7 else None for i in range(n)]
    metas = [{"document_path": f"p{i % 5}.html", "chunk_index": i, "score": 0.5 * i, "is_enterprise": bool(i % 2)}
             if i % 9 else None for i in range(n)]
    for as_object in (False, True):
        d = tmp_path / f"store_{int(as_object)}"
        info = cf.write_store(str(d), ids, docs, metas, emb, n_flushed=40, deleted_labels=(999, 1000), pickle_as_object=as_object)
        st = cs.read_chroma_store(str(d))
        assert st["ids"] == ids and st["documents"] == docs and st["metadatas"] == metas
        assert st["dim"] == dim and st["metadata"] == {"hnsw:space": "cosine"}
        assert np.array_equal(st["embeddings"][:40], info["normalized_flushed"].astype(np.float32))
        assert np.array_equal(st["embeddings"][40:], emb[40:])
    with pytest.raises(KeyError):
        cs.read_chroma_store(str(d), "no_such_collection")


def test_bench_reads_roofline_traffic_from_the_committed_profiles():
    """bench.py's `roofline.traffic` is the DRAM byte count of the named kernel in the committed ncu raw pages (never a
    literal): the reader finds both kernels and returns bytes in a plausible range, and None for what is not there"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    dense = bench.profile_traffic("r2_dense_step_ncu_full_raw.csv", "dense_gemm_topk_kernel<1, 2>")
    assert dense is not None and 2.0e9 < dense["bytes_per_launch"] < 2.3e9       # one pass over 1M x 1024 bf16 + lists
    bm25 = bench.profile_traffic("r2_bm25_filter_ncu_full_raw.csv", "bm25_filter")
    assert bm25 is not None and 1.0e8 < bm25["bytes_per_launch"] < 2.9e9          # <= the algorithmic 2.78 GB
    assert bench.profile_traffic("r2_bm25_filter_ncu_full_raw.csv", "no_such_kernel") is None
    assert bench.profile_traffic("no_such_file.csv", "bm25_filter") is None
