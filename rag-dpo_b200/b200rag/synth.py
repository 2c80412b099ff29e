"""Synthetic inputs for benchmarks and tests (numpy twins of the device-side
generators), shaped like the workloads in BASELINE.json."""
import numpy as np

_G1 = np.uint64(0x9E3779B97F4A7C15)
_G2 = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def synth_values(seed, row0, nrows, dim):
    """un-normalised values of rag_corpus_fill_synthetic (csrc/corpus.cu synth_value)"""
    with np.errstate(over="ignore"):
        r = (np.arange(row0, row0 + nrows, dtype=np.uint64) * _G1)[:, None]
        c = (np.arange(dim, dtype=np.uint64) * _G2)[None, :]
        z = np.uint64(seed) + r + c
        z ^= z >> np.uint64(30); z *= _M1
        z ^= z >> np.uint64(27); z *= _M2
        z ^= z >> np.uint64(31)
    u = (z >> np.uint64(40)).astype(np.float32)
    return u * np.float32(1.0 / 8388608.0) - np.float32(1.0)


def synth_rows(seed, row0, nrows, dim, chunk=65536):
    """fp32 rows before the storage-dtype rounding: canonical-order fp64 norm, fp64 divide"""
    if nrows > chunk:
        out = np.empty((nrows, dim), dtype=np.float32)
        for s in range(0, nrows, chunk):
            n = min(chunk, nrows - s)
            out[s:s + n] = synth_rows(seed, row0 + s, n, dim, chunk)
        return out
    v = synth_values(seed, row0, nrows, dim).astype(np.float64)
    p = np.zeros((nrows, 32), dtype=np.float64)
    for j in range(dim // 32):
        blk = v[:, 32 * j:32 * j + 32]
        p += blk * blk
    off = 16
    while off >= 1:
        p[:, :off] += p[:, off:2 * off]
        off //= 2
    return (v / np.sqrt(p[:, :1])).astype(np.float32)


def unit_queries(n, dim, seed):
    """fp32 unit query vectors (standard normal, L2-normalised)"""
    g = np.random.default_rng(seed)
    q = g.standard_normal((n, dim)).astype(np.float32)
    q64 = q.astype(np.float64)
    return (q64 / np.sqrt((q64 ** 2).sum(axis=1, keepdims=True))).astype(np.float32)


def zipf_corpus(n_docs, vocab, seed, lo=40, hi=250, s=1.07):
    """Tokenised synthetic corpus for the keyword leg: Zipf(s) term ids, doc length U[lo,hi],
    relabelled so that term ids are in first-seen order.  Returns (list of int arrays, n_terms)."""
    g = np.random.default_rng(seed)
    lens = g.integers(lo, hi + 1, size=n_docs)
    p = np.arange(1, vocab + 1, dtype=np.float64) ** (-s)
    p /= p.sum()
    flat = g.choice(vocab, size=int(lens.sum()), p=p)
    uniq, first = np.unique(flat, return_index=True)
    remap = np.full(vocab, -1, dtype=np.int64)
    remap[uniq[np.argsort(first)]] = np.arange(len(uniq))
    flat = remap[flat]
    return np.split(flat, np.cumsum(lens)[:-1]), len(uniq)
