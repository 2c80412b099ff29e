"""Loader for the reference's on-disk vector store (data/vectordb/chromadb: `chroma.sqlite3` + one HNSW segment
directory per collection) into a DeviceCollection, without the chromadb wheel.

The reference writes the store with chromadb.PersistentClient / collection.add
(src/processing/create_chromadb_index.py:100-130, 374-379; src/processing/ingest_enterprise.py:241-246), ships it
as a zip of exactly these files (scripts/package_cnil_db.py:33-58) and opens it again in app.py:42-119.

FORMAT STATUS: PARITY UNPINNED.  chromadb==1.4.1 is neither vendored nor installable here and the reference holds
no copy of a store, so the layout below is restated from the published chromadb / chroma-hnswlib sources
(local persistent segments) and exercised against a fixture written by oracle/chroma_fixture.py in the same
layout; the first contact with a real store must re-verify it.

    chroma.sqlite3
        collections(id, name, dimension, ...)
        segments(id, type, scope, collection)            scope 'VECTOR' -> directory <id>/, scope 'METADATA'
        embeddings(id INTEGER, segment_id, embedding_id, seq_id, ...)      embedding_id = the caller's string id
        embedding_metadata(id, key, string_value, int_value, float_value, bool_value)   'chroma:document' = the text
        embeddings_queue(seq_id, operation, topic, id, vector BLOB float32, encoding, metadata)   the write-ahead log
    <vector segment id>/
        header.bin          int32 version | offsetLevel0, max_elements, cur_element_count, size_data_per_element,
                            label_offset, offsetData (u64 each) | int32 maxlevel | u32 enterpoint | maxM, maxM0, M (u64)
                            | double mult | u64 ef_construction                     (100 bytes)
        data_level0.bin     cur_element_count x size_data_per_element: [level-0 links | vector fp32 x dim | label u64]
        index_metadata.pickle   id_to_label / label_to_id / max_seq_id / dimensionality
Cosine space: hnswlib stores the vectors already L2-normalised; records that only live in the write-ahead log
(not yet flushed into the index: chromadb's sync threshold) are raw.  DeviceCollection.add normalises either way.
"""
import io
import os
import pickle
import sqlite3
import struct

import numpy as np

_HEADER = struct.Struct("<i6QiI3QdQ")          # 100 bytes
_DELETE_MARK = 0x01                             # hnswlib: bit 0 of byte 2 of an element's level-0 link header

OP_ADD, OP_UPDATE, OP_UPSERT, OP_DELETE = 0, 1, 2, 3


class _Bag:
    """stands in for chromadb's PersistentData class when unpickling index_metadata.pickle"""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


class _LenientUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("chromadb"):
            return _Bag
        return super().find_class(module, name)


def _index_metadata(path):
    with open(path, "rb") as f:
        obj = _LenientUnpickler(io.BytesIO(f.read())).load()
    d = obj if isinstance(obj, dict) else obj.__dict__
    return {"id_to_label": dict(d.get("id_to_label", {})), "max_seq_id": d.get("max_seq_id", 0),
            "dimensionality": d.get("dimensionality")}


def _seq_to_int(v):
    if v is None:
        return 0
    if isinstance(v, (bytes, bytearray, memoryview)):
        return int.from_bytes(bytes(v), "big")
    return int(v)


def read_hnsw_vectors(segment_dir, dim=None):
    """-> {label: fp32 vector} of the live elements of a persisted hnswlib index (level-0 data only)"""
    with open(os.path.join(segment_dir, "header.bin"), "rb") as f:
        raw = f.read()
    if len(raw) < _HEADER.size:
        raise ValueError(f"{segment_dir}/header.bin: {len(raw)} bytes, expected {_HEADER.size}")
    (_version, off_l0, _max_el, count, per_el, label_off, off_data, _maxlevel, _enter, _maxm, _maxm0, _m, _mult,
     _efc) = _HEADER.unpack_from(raw)
    data_bytes = label_off - off_data
    if data_bytes <= 0 or data_bytes % 4 or per_el < label_off + 8:
        raise ValueError(f"{segment_dir}/header.bin: inconsistent offsets")
    if dim is not None and data_bytes != 4 * dim:
        raise ValueError(f"{segment_dir}: vectors of {data_bytes // 4} floats, collection dimension {dim}")
    blob = np.fromfile(os.path.join(segment_dir, "data_level0.bin"), dtype=np.uint8, count=count * per_el)
    if blob.size != count * per_el:
        raise ValueError(f"{segment_dir}/data_level0.bin: truncated")
    el = blob.reshape(count, per_el)
    live = (el[:, off_l0 + 2] & _DELETE_MARK) == 0
    vec = np.ascontiguousarray(el[:, off_data:off_data + data_bytes]).view(np.float32)
    labels = np.ascontiguousarray(el[:, label_off:label_off + 8]).view(np.uint64).ravel()
    return {int(l): vec[i] for i, l in enumerate(labels) if live[i]}


def read_chroma_store(path, collection_name="rag_dpo_chunks"):
    """-> dict(ids, documents, metadatas, embeddings (n, dim) fp32, metadata of the collection): the rows of the
    collection in insertion order (embeddings.id), i.e. what collection.get(include=[documents, metadatas,
    embeddings]) returns"""
    db = sqlite3.connect(f"file:{os.path.join(path, 'chroma.sqlite3')}?mode=ro", uri=True)
    try:
        row = db.execute("SELECT id, dimension FROM collections WHERE name = ?", (collection_name,)).fetchone()
        if row is None:
            raise KeyError(f"collection {collection_name!r} not in {path}")
        cid, dim = row
        segs = {scope: sid for sid, scope in db.execute("SELECT id, scope FROM segments WHERE collection = ?", (cid,))}
        meta_seg, vec_seg = segs.get("METADATA"), segs.get("VECTOR")
        cmeta = {}
        try:
            for key, sv, iv, fv in db.execute("SELECT key, str_value, int_value, float_value FROM collection_metadata "
                                              "WHERE collection_id = ?", (cid,)):
                cmeta[key] = sv if sv is not None else (iv if iv is not None else fv)
        except sqlite3.Error:
            pass
        rows = db.execute("SELECT id, embedding_id FROM embeddings WHERE segment_id = ? ORDER BY id", (meta_seg,)).fetchall()
        ids = [r[1] for r in rows]
        pos = {r[0]: i for i, r in enumerate(rows)}
        documents = [None] * len(ids)
        metadatas = [None] * len(ids)
        q = ("SELECT m.id, m.key, m.string_value, m.int_value, m.float_value, m.bool_value FROM embedding_metadata m "
             "JOIN embeddings e ON e.id = m.id WHERE e.segment_id = ?")
        for rid, key, sv, iv, fv, bv in db.execute(q, (meta_seg,)):
            i = pos[rid]
            if key == "chroma:document":
                documents[i] = sv
                continue
            val = sv if sv is not None else (bool(bv) if bv is not None else (iv if iv is not None else fv))
            if metadatas[i] is None:
                metadatas[i] = {}
            metadatas[i][key] = val
        # ---- vectors: the flushed part from the HNSW segment, the tail from the write-ahead log
        vectors = {}
        max_seq = 0
        seg_dir = os.path.join(path, vec_seg) if vec_seg else None
        if seg_dir and os.path.isfile(os.path.join(seg_dir, "header.bin")):
            im = _index_metadata(os.path.join(seg_dir, "index_metadata.pickle"))
            max_seq = _seq_to_int(im["max_seq_id"])
            by_label = read_hnsw_vectors(seg_dir, dim)
            for sid, label in im["id_to_label"].items():
                v = by_label.get(int(label))
                if v is not None:
                    vectors[sid] = v
        wal = db.execute("SELECT seq_id, operation, id, vector, encoding FROM embeddings_queue WHERE topic LIKE ? "
                         "ORDER BY seq_id", (f"%{cid}",)).fetchall()
        for seq, op, sid, blob, enc in wal:
            if _seq_to_int(seq) <= max_seq:
                continue
            if op == OP_DELETE:
                vectors.pop(sid, None)
            elif blob is not None:
                if enc not in (None, "FLOAT32"):
                    raise ValueError(f"embeddings_queue: encoding {enc!r} not supported")
                vectors[sid] = np.frombuffer(blob, dtype="<f4")
    finally:
        db.close()
    missing = [i for i in ids if i not in vectors]
    if missing:
        raise ValueError(f"{len(missing)} ids have no vector in the segment or the log (first: {missing[0]!r})")
    if dim is None and ids:
        dim = len(vectors[ids[0]])
    emb = np.stack([vectors[i] for i in ids]).astype(np.float32) if ids else np.zeros((0, dim or 0), np.float32)
    return {"ids": ids, "documents": documents, "metadatas": metadatas, "embeddings": emb, "dim": dim, "metadata": cmeta}


def load_collection(path, collection_name="rag_dpo_chunks", dtype="f32", batch=20000, **kwargs):
    """the collection of a chromadb persist directory as a DeviceCollection (what app.py:42-119 obtains with
    chromadb.PersistentClient(path).get_collection(name))"""
    from .collection import DeviceCollection
    st = read_chroma_store(path, collection_name)
    col = DeviceCollection(name=collection_name, dim=st["dim"], dtype=dtype, metadata=st["metadata"] or None,
                           capacity=len(st["ids"]), **kwargs)
    for a in range(0, len(st["ids"]), batch):
        b = min(a + batch, len(st["ids"]))
        col.add(ids=st["ids"][a:b], documents=st["documents"][a:b], embeddings=st["embeddings"][a:b],
                metadatas=st["metadatas"][a:b])
    return col
