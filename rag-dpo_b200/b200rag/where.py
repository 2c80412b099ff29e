"""`where` metadata predicates.

The reference passes Chroma `where` dicts to collection.query
(src/rag/retriever.py:215-220, 380-385); shapes in use: {"field": v},
{"field": {"$ne": v}}, {"field": {"$in": [...]}}, {"$and": [...]},
{"$or": [...]} (src/rag/pipeline.py:59-69, pages/1_*Chat.py:247,
test_rag.py:145, src/processing/ingest_enterprise.py:291-294).  Chroma filters
BEFORE the kNN, so the kernels test an allowed-row bitmap ahead of the top-k.

Two forms:
  * ColumnCodes + compile_where: metadata values are dictionary-coded into int32
    columns that live on the device (rag_corpus_set_codes); a `where` dict
    compiles to a small postfix program (csrc/rowfilter.cu) that the device
    evaluates into the bitmap — nothing per row happens on the host at query time.
  * match(): the host evaluator, for collection.get / collection.delete(where=)
    and for operators the program does not cover ($gt/$gte/$lt/$lte).
"""
import numpy as np

OP_TRUE, OP_FALSE, OP_EQ, OP_NE, OP_IN, OP_NIN, OP_AND, OP_OR = range(8)
MAX_COLUMNS = 32
NO_SUCH_VALUE = -2           # a value no row carries: EQ never matches, NE always does


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


class Unsupported(Exception):
    """the predicate needs the host evaluator"""


class ColumnCodes:
    """metadata key -> device column, and per column value -> int32 code.  Values are keyed by (type, value) so
    that True and 1 stay different (Chroma compares typed values)."""

    def __init__(self):
        self.column = {}            # key -> column index
        self.codes = []             # per column: {(type name, value): code}
        self.uncoded = set()        # keys beyond MAX_COLUMNS or with unhashable / non-scalar values

    def _key(self, value):
        return (type(value).__name__, value)

    def encode_batch(self, metadatas):
        """codes of a batch of metadata dicts: {column index: int32 array (len(metadatas)), -1 = key missing};
        only columns some row of the batch carries are returned"""
        out = {}
        n = len(metadatas)
        for j, m in enumerate(metadatas):
            if not m:
                continue
            for key, value in m.items():
                if key in self.uncoded:
                    continue
                col = self.column.get(key)
                if col is None:
                    if len(self.codes) >= MAX_COLUMNS or not isinstance(value, (str, bool, int, float)):
                        self.uncoded.add(key)
                        continue
                    col = self.column[key] = len(self.codes)
                    self.codes.append({})
                if not isinstance(value, (str, bool, int, float)):
                    # a non-scalar value under a coded key can never equal a scalar operand: leave it "missing"
                    continue
                table = self.codes[col]
                k = self._key(value)
                code = table.get(k)
                if code is None:
                    code = table[k] = len(table)
                arr = out.get(col)
                if arr is None:
                    arr = out[col] = np.full(n, -1, dtype=np.int32)
                arr[j] = code
        return out

    def code_of(self, key, value):
        """(column, code) for an operand; column None = no row has the key"""
        if key in self.uncoded:
            raise Unsupported(key)
        col = self.column.get(key)
        if col is None:
            return None, NO_SUCH_VALUE
        try:
            return col, self.codes[col].get(self._key(value), NO_SUCH_VALUE)
        except TypeError:           # unhashable operand
            return col, NO_SUCH_VALUE


def compile_where(where, cols, doc_paths=None, doc_key="document_path"):
    """where dict (+ optional set of document paths: an IN on `doc_key`) -> int32 postfix program for the device
    (csrc/rowfilter.cu), or None when there is nothing to filter.  Raises Unsupported for what only the host
    evaluator covers."""

    def leaf(op, key, val):
        if op in ("$eq", "$ne"):
            col, code = cols.code_of(key, val)
            if col is None:
                return [OP_FALSE if op == "$eq" else OP_TRUE]
            return [OP_EQ if op == "$eq" else OP_NE, col, code]
        if op in ("$in", "$nin"):
            col = None
            present = []
            for x in val:
                c, code = cols.code_of(key, x)
                col = c if c is not None else col
                if code >= 0:
                    present.append(code)
            if col is None or not present:
                return [OP_FALSE if op == "$in" else OP_TRUE]
            nbits = max(present) + 1
            words = np.zeros((nbits + 31) // 32, dtype=np.uint32)
            for code in present:
                words[code >> 5] |= np.uint32(1) << np.uint32(code & 31)
            return [OP_IN if op == "$in" else OP_NIN, col, nbits, len(words)] + words.view(np.int32).tolist()
        raise Unsupported(op)

    def fold(parts, op):
        if len(parts) == 1:
            return parts[0]
        if len(parts) > 32:
            raise Unsupported("more than 32 operands")
        return [w for p in parts for w in p] + [op, len(parts)]

    def node(w):
        parts = []
        for key, cond in w.items():
            if key == "$and":
                parts.append(fold([node(x) for x in cond], OP_AND) if cond else [OP_TRUE])
            elif key == "$or":
                parts.append(fold([node(x) for x in cond], OP_OR) if cond else [OP_FALSE])
            elif isinstance(cond, dict):
                sub = [leaf(op, key, val) for op, val in cond.items()]
                parts.append(fold(sub, OP_AND) if sub else [OP_TRUE])
            else:
                parts.append(leaf("$eq", key, cond))
        return fold(parts, OP_AND) if parts else [OP_TRUE]

    parts = []
    if where:
        parts.append(node(where))
    if doc_paths is not None:
        parts.append(leaf("$in", doc_key, list(doc_paths)))
    if not parts:
        return None
    return np.asarray(fold(parts, OP_AND), dtype=np.int32)


def run_program(prog, codes_of_row):
    """host interpreter of the postfix program (tests): codes_of_row(col) -> code of the row (-1 missing)"""
    stack = []
    pc = 0
    prog = [int(x) for x in prog]
    while pc < len(prog):
        op = prog[pc]
        if op == OP_TRUE:
            stack.append(True); pc += 1
        elif op == OP_FALSE:
            stack.append(False); pc += 1
        elif op in (OP_EQ, OP_NE):
            c = codes_of_row(prog[pc + 1])
            stack.append((c == prog[pc + 2]) == (op == OP_EQ)); pc += 3
        elif op in (OP_IN, OP_NIN):
            c = codes_of_row(prog[pc + 1])
            nbits, nwords = prog[pc + 2], prog[pc + 3]
            inside = 0 <= c < nbits and ((prog[pc + 4 + (c >> 5)] & 0xFFFFFFFF) >> (c & 31)) & 1 == 1
            stack.append(inside == (op == OP_IN)); pc += 4 + nwords
        else:
            n = prog[pc + 1]
            top = stack[len(stack) - n:]
            del stack[len(stack) - n:]
            stack.append(all(top) if op == OP_AND else any(top)); pc += 2
    return stack[-1] if stack else True
