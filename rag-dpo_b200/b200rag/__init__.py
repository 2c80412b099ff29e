"""b200rag — B200-native (sm_100a) retrieval hot path behind RAG-DPO's own
injection points (src/rag/retriever.py:107-148): dense similarity + exact
top-k, BM25 keyword scoring over CSR postings, weighted RRF.

The compute lives in libb200rag.so (C ABI: include/b200rag.h); this package is
the host-side mirror of the reference's operator interface.  No CPU fallback:
the device classes raise if the library or an sm_100 GPU is missing.
"""
from ._lib import B200RagError, RAG_BF16, RAG_F16, RAG_F32, RAG_MAX_K, pinned_empty
from .bm25 import BM25Result, DeviceBM25, DeviceChunkBM25Index, DeviceSummaryBM25Index, Postings
from .collection import DeviceCollection, DeviceCorpus, distance_from_score, l2_normalize_rows
from .reranker import DeviceRerankStep, RankedChunk, rerank_select
from .retriever import HybridRetriever, RetrievedChunk, RetrievedDocument
from .rrf import fuse_ranked, reciprocal_rank_fusion, rrf_fuse_rows
from .tokenizer import tokenize_french

__all__ = [
    "B200RagError", "pinned_empty", "RAG_F32", "RAG_BF16", "RAG_F16", "RAG_MAX_K",
    "DeviceCollection", "DeviceCorpus", "l2_normalize_rows", "distance_from_score",
    "DeviceChunkBM25Index", "DeviceSummaryBM25Index", "DeviceBM25", "Postings", "BM25Result",
    "HybridRetriever", "RetrievedChunk", "RetrievedDocument",
    "reciprocal_rank_fusion", "fuse_ranked", "rrf_fuse_rows", "tokenize_french",
    "DeviceRerankStep", "RankedChunk", "rerank_select",
]
