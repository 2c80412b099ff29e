"""ORACLE — test infrastructure, not product code.

Writes a directory in the layout chromadb's local persistent client leaves on disk (chroma.sqlite3 + an HNSW
segment directory), restated from the published chromadb / chroma-hnswlib sources — the wheel is absent, so this
fixture is PARITY UNPINNED like the loader it exercises (rag-dpo_b200/b200rag/chroma_store.py).  The HNSW graph
itself (link lists) is left empty: the loader reads level-0 data only."""
import os
import pickle
import sqlite3
import struct
import sys
import types
import uuid

import numpy as np

_HEADER = struct.Struct("<i6QiI3QdQ")


def _typed(v):
    """(string_value, int_value, float_value, bool_value) of one metadata value"""
    if isinstance(v, bool):
        return (None, None, None, int(v))
    if isinstance(v, int):
        return (None, v, None, None)
    if isinstance(v, float):
        return (None, None, v, None)
    return (str(v), None, None, None)


def write_store(path, ids, documents, metadatas, embeddings, n_flushed, deleted_labels=(), name="rag_dpo_chunks",
                pickle_as_object=False, M=16):
    """rows [0, n_flushed) live in the HNSW segment (L2-normalised, cosine space), the rest only in the write-ahead
    log (raw).  deleted_labels: extra elements written into the segment and marked deleted (must be ignored)."""
    os.makedirs(path, exist_ok=True)
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    n, dim = emb.shape
    cid, meta_seg, vec_seg = (str(uuid.UUID(int=i + 1)) for i in range(3))
    db = sqlite3.connect(os.path.join(path, "chroma.sqlite3"))
    db.executescript("""
        CREATE TABLE collections (id TEXT PRIMARY KEY, name TEXT NOT NULL, dimension INTEGER, database_id TEXT, config_json_str TEXT);
        CREATE TABLE collection_metadata (collection_id TEXT, key TEXT, str_value TEXT, int_value INTEGER, float_value REAL);
        CREATE TABLE segments (id TEXT PRIMARY KEY, type TEXT NOT NULL, scope TEXT NOT NULL, collection TEXT);
        CREATE TABLE embeddings (id INTEGER PRIMARY KEY, segment_id TEXT NOT NULL, embedding_id TEXT NOT NULL, seq_id BLOB NOT NULL,
                                 created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP);
        CREATE TABLE embedding_metadata (id INTEGER, key TEXT NOT NULL, string_value TEXT, int_value INTEGER, float_value REAL,
                                         bool_value INTEGER, PRIMARY KEY (id, key));
        CREATE TABLE embeddings_queue (seq_id INTEGER PRIMARY KEY, created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP, operation INTEGER NOT NULL,
                                       topic TEXT NOT NULL, id TEXT NOT NULL, vector BLOB, encoding TEXT, metadata TEXT);
    """)
    db.execute("INSERT INTO collections VALUES (?,?,?,?,?)", (cid, name, dim, "default", "{}"))
    db.execute("INSERT INTO collection_metadata VALUES (?,?,?,?,?)", (cid, "hnsw:space", "cosine", None, None))
    db.execute("INSERT INTO segments VALUES (?,?,?,?)", (vec_seg, "urn:chroma:segment/vector/hnsw-local-persisted", "VECTOR", cid))
    db.execute("INSERT INTO segments VALUES (?,?,?,?)", (meta_seg, "urn:chroma:segment/metadata/sqlite", "METADATA", cid))
    topic = f"persistent://default/default/{cid}"
    for i in range(n):
        seq = i + 1
        db.execute("INSERT INTO embeddings (id, segment_id, embedding_id, seq_id) VALUES (?,?,?,?)",
                   (i + 1, meta_seg, ids[i], seq.to_bytes(8, "big")))
        if documents[i] is not None:
            db.execute("INSERT INTO embedding_metadata VALUES (?,?,?,?,?,?)", (i + 1, "chroma:document", documents[i], None, None, None))
        for k, v in (metadatas[i] or {}).items():
            db.execute("INSERT INTO embedding_metadata VALUES (?,?,?,?,?,?)", (i + 1, k) + _typed(v))
        if i >= n_flushed:
            db.execute("INSERT INTO embeddings_queue (seq_id, operation, topic, id, vector, encoding, metadata) VALUES (?,?,?,?,?,?,?)",
                       (seq, 0, topic, ids[i], emb[i].astype("<f4").tobytes(), "FLOAT32", None))
    # a stale log record below the index's max_seq_id must be ignored
    if n_flushed > 0:
        db.execute("INSERT INTO embeddings_queue (seq_id, operation, topic, id, vector, encoding, metadata) VALUES (?,?,?,?,?,?,?)",
                   (n + 100 - n - 100 + 0 - 0 or -1, 0, topic, ids[0], np.full(dim, 7.0, "<f4").tobytes(), "FLOAT32", None))
    db.commit()
    db.close()
    # ---- HNSW segment
    seg = os.path.join(path, vec_seg)
    os.makedirs(seg, exist_ok=True)
    maxM0 = 2 * M
    size_links0 = maxM0 * 4 + 4
    off_data, label_off = size_links0, size_links0 + dim * 4
    per_el = label_off + 8
    labels = list(range(1, n_flushed + 1)) + list(deleted_labels)
    count = len(labels)
    el = np.zeros((count, per_el), dtype=np.uint8)
    norm = emb[:n_flushed] / np.maximum(np.linalg.norm(emb[:n_flushed].astype(np.float64), axis=1, keepdims=True), 1e-30)
    for j, lab in enumerate(labels):
        v = norm[j].astype("<f4") if j < n_flushed else np.full(dim, 3.0, "<f4")
        el[j, off_data:label_off] = np.frombuffer(v.tobytes(), dtype=np.uint8)
        el[j, label_off:] = np.frombuffer(struct.pack("<Q", lab), dtype=np.uint8)
        if j >= n_flushed:
            el[j, 2] |= 0x01                      # DELETE_MARK
    el.tofile(os.path.join(seg, "data_level0.bin"))
    with open(os.path.join(seg, "header.bin"), "wb") as f:
        f.write(_HEADER.pack(1, 0, max(count, 1000), count, per_el, label_off, off_data, 0, 0, M, maxM0, M, 1.0 / np.log(M), 100))
    np.zeros(count, np.int32).tofile(os.path.join(seg, "length.bin"))
    open(os.path.join(seg, "link_lists.bin"), "wb").close()
    state = {"dimensionality": dim, "total_elements_added": count, "max_seq_id": n_flushed,
             "id_to_label": {ids[i]: i + 1 for i in range(n_flushed)}, "label_to_id": {i + 1: ids[i] for i in range(n_flushed)},
             "id_to_seq_id": {ids[i]: i + 1 for i in range(n_flushed)}}
    with open(os.path.join(seg, "index_metadata.pickle"), "wb") as f:
        if pickle_as_object:
            # older chromadb pickles a PersistentData instance by reference to its module
            modname = "chromadb.segment.impl.vector.local_persistent_hnsw"
            created = []
            parts = modname.split(".")
            for i in range(1, len(parts) + 1):
                mn = ".".join(parts[:i])
                if mn not in sys.modules:
                    sys.modules[mn] = types.ModuleType(mn)
                    created.append(mn)
            cls = type("PersistentData", (), {"__module__": modname})
            setattr(sys.modules[modname], "PersistentData", cls)
            obj = cls()
            obj.__dict__.update(state)
            try:
                pickle.dump(obj, f)
            finally:
                for mn in created:
                    sys.modules.pop(mn, None)
        else:
            pickle.dump(state, f)
    return {"collection_id": cid, "vector_segment": vec_seg, "normalized_flushed": norm}
