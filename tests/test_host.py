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
    assert L.rag_abi_version() == 1


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
