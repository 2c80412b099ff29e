"""`where` metadata predicate -> allowed-row bitmap (host side).

The reference passes Chroma `where` dicts to collection.query
(src/rag/retriever.py:215-220, 380-385); shapes in use: {"field": v},
{"field": {"$ne": v}}, {"field": {"$in": [...]}}, {"$and": [...]},
{"$or": [...]} (src/rag/pipeline.py:59-69, pages/1_*Chat.py:247,
test_rag.py:145, src/processing/ingest_enterprise.py:291-294).  Chroma filters
BEFORE the kNN, so the kernels take the predicate as a row bitmap and test it
ahead of the top-k insert.
"""
import json

import numpy as np


def _same(a, b):
    return type(a) is type(b) and a == b


def match(meta, where):
    if not where:
        return True
    meta = meta or {}
    for key, cond in where.items():
        if key == "$and":
            ok = all(match(meta, w) for w in cond)
        elif key == "$or":
            ok = any(match(meta, w) for w in cond)
        elif isinstance(cond, dict):
            ok = True
            for op, val in cond.items():
                has = key in meta
                v = meta.get(key)
                if op == "$eq":
                    r = has and _same(v, val)
                elif op == "$ne":
                    r = not (has and _same(v, val))
                elif op == "$in":
                    r = has and any(_same(v, x) for x in val)
                elif op == "$nin":
                    r = not (has and any(_same(v, x) for x in val))
                elif op in ("$gt", "$gte", "$lt", "$lte"):
                    if not has or isinstance(v, (str, bool)) or isinstance(val, (str, bool)):
                        r = False
                    else:
                        r = {"$gt": v > val, "$gte": v >= val, "$lt": v < val, "$lte": v <= val}[op]
                else:
                    raise ValueError(f"unsupported where operator: {op}")
                ok = ok and r
        else:
            ok = key in meta and _same(meta[key], cond)
        if not ok:
            return False
    return True


def bitmap_from_mask(mask):
    """bool array (n,) -> uint8 bitmap, bit r of byte r//8 (little bit order)."""
    return np.packbits(np.asarray(mask, dtype=bool), bitorder="little")


class WhereCompiler:
    """Caches compiled bitmaps per (filter, collection version)."""

    def __init__(self):
        self._cache = {}
        self._version = 0

    def invalidate(self):
        self._version += 1
        self._cache.clear()

    def compile(self, metadatas, where):
        if not where:
            return None, len(metadatas)
        key = json.dumps(where, sort_keys=True, default=str)
        hit = self._cache.get(key)
        if hit is None:
            mask = np.fromiter((match(m, where) for m in metadatas), dtype=bool, count=len(metadatas))
            hit = (bitmap_from_mask(mask), int(mask.sum()))
            if len(self._cache) > 64:
                self._cache.clear()
            self._cache[key] = hit
        return hit
