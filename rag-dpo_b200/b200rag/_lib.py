"""ctypes binding of libb200rag.so (C ABI declared in include/b200rag.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 GPU
can be bound, every compute entry point raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rag.so")

RAG_F32, RAG_BF16, RAG_F16 = 0, 1, 2
RAG_MAX_K = 224
DTYPES = {"f32": RAG_F32, "fp32": RAG_F32, "float32": RAG_F32, "bf16": RAG_BF16, "bfloat16": RAG_BF16,
          "f16": RAG_F16, "fp16": RAG_F16, "float16": RAG_F16}

# every symbol include/b200rag.h declares: (restype, argtypes)
_vp, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
SYMBOLS = {
    "rag_init": (_i, [_i]),
    "rag_init_devices": (_i, [_i, _vp]),
    "rag_slot_count": (_i, [_vp]),
    "rag_set_stream": (_i, [_vp]),
    "rag_last_error": (C.c_char_p, []),
    "rag_abi_version": (_i, []),
    "rag_device_info": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "rag_set_option": (_i, [C.c_char_p, _i64]),
    "rag_host_alloc": (_i, [_vp, C.c_size_t]),
    "rag_host_free": (_i, [_vp]),
    "rag_last_timings": (_i, [_vp, _i]),
    "rag_counters": (_i, [_vp, _i]),
    "rag_debug_last_candidates": (_i, [_vp, _i, _vp, _i64, _vp, _vp]),
    "rag_corpus_create": (_i, [_vp, _i64, _i, _i]),
    "rag_corpus_create_sharded": (_i, [_vp, _i64, _i, _i, _i]),
    "rag_corpus_delete_rows": (_i, [_vp, _vp, _i64]),
    "rag_corpus_live_count": (_i, [_vp, _vp]),
    "rag_corpus_set_codes": (_i, [_vp, _i, _i64, _i64, _vp]),
    "rag_dense_topk_where": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "rag_corpus_destroy": (_i, [_vp]),
    "rag_corpus_reserve": (_i, [_vp, _i64]),
    "rag_corpus_upload": (_i, [_vp, _i64, _i64, _vp]),
    "rag_corpus_download": (_i, [_vp, _i64, _i64, _vp]),
    "rag_corpus_compact": (_i, [_vp, _vp, _i64]),
    "rag_corpus_count": (_i, [_vp, _vp]),
    "rag_corpus_fill_synthetic": (_i, [_vp, _u64, _i64, _i64, _i64]),
    "rag_corpus_device_ptr": (_i, [_vp, _vp]),
    "rag_dense_topk": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "rag_dense_topk_dev": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "rag_merge_topk_dev": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp]),
    "rag_exchange_create": (_i, [_vp, _i, _i, C.c_size_t, _vp]),
    "rag_exchange_connect": (_i, [_vp, _vp]),
    "rag_exchange_destroy": (_i, [_vp]),
    "rag_exchange_merge_topk_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "rag_exchange_merge_rows_dev": (_i, [_vp, _vp, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp]),
    "rag_exchange_status": (_i, [_vp, _vp]),
    "rag_csr_build": (_i, [_i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i]),
    "rag_bm25_create": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _d, _d, _d]),
    "rag_bm25_create_sharded": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _d, _d, _d]),
    "rag_bm25_destroy": (_i, [_vp]),
    "rag_bm25_info": (_i, [_vp, _vp, _vp]),
    "rag_bm25_query_bytes": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "rag_bm25_search": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "rag_bm25_scores": (_i, [_vp, _vp, _i, _vp]),
    "rag_rrf_fuse": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rag_rerank_select": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp]),
}

_lib = None
_ready = False
_lock = threading.Lock()


class B200RagError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every symbol (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200RagError(
                f"{LIB_PATH} is missing: build it with `make -C rag-dpo_b200/csrc` "
                "(or __graft_entry__.build()); b200rag has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B200RagError(f"libb200rag error {rc}: {load().rag_last_error().decode(errors='replace')}")


def lib():
    """The library bound to this process's primary GPU (B200RAG_DEVICE, else LOCAL_RANK, default 0)."""
    global _ready
    L = load()
    if not _ready:
        with _lock:
            if not _ready:
                dev = int(os.environ.get("B200RAG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
                check(L.rag_init(dev))
                _ready = True
    return L


def init_devices(devices):
    """One process, several GPUs: shard slot i of a sharded corpus / index runs on devices[i] (a device may repeat).
    devices[0] must be the primary device lib() binds.  Idempotent; a longer list extends a shorter one."""
    L = lib()
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    check(L.rag_init_devices(len(devices), arr))
    return slot_count()


def slot_count():
    n = C.c_int()
    check(lib().rag_slot_count(C.byref(n)))
    return n.value


def ensure_slots(n_shards, devices=None):
    """make at least n_shards shard slots available: the given devices, else B200RAG_DEVICES ("0,1,2,3"), else the
    primary device followed by the other visible GPUs in order (wrapping around when there are fewer GPUs)"""
    if n_shards <= slot_count() and devices is None:
        return
    if devices is None:
        env = os.environ.get("B200RAG_DEVICES")
        if env:
            devices = [int(x) for x in env.split(",") if x.strip() != ""]
        else:
            import torch
            n_gpu = max(1, torch.cuda.device_count())
            first = int(os.environ.get("B200RAG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            devices = [(first + i) % n_gpu for i in range(n_shards)]
    if len(devices) < n_shards:
        raise B200RagError(f"{n_shards} shards need {n_shards} shard slots, got devices {devices}")
    init_devices(devices)


def ptr(a):
    """address of a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class _PinnedBlock:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().rag_host_alloc(C.byref(self.ptr), max(int(nbytes), 1)))

    def __del__(self):
        try:
            if self.ptr:
                load().rag_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory: the host-pointer entry points DMA it directly."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    block = _PinnedBlock(n)
    buf = (C.c_uint8 * max(n, 1)).from_address(block.ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED_KEEPALIVE[id(buf)] = block
    import weakref
    weakref.finalize(buf, _PINNED_KEEPALIVE.pop, id(buf), None)
    return arr


_PINNED_KEEPALIVE = {}


def device_info():
    L = lib()
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    fr, to = C.c_size_t(), C.c_size_t()
    check(L.rag_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(fr), C.byref(to)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "free_bytes": fr.value, "total_bytes": to.value}


def last_timings():
    ms = np.zeros(8, dtype=np.float32)
    check(lib().rag_last_timings(ptr(ms), 8))
    return ms


def counters():
    out = np.zeros(4, dtype=np.int64)
    check(lib().rag_counters(ptr(out), 4))
    return {"launches": int(out[0]), "fallbacks": int(out[1]), "fallback_queries": int(out[2])}


def sync_stream_of(torch, device):
    """rag_merge_topk_dev is stream-ordered: wait for the library's stream before reading its output"""
    torch.cuda.synchronize(device)


def set_option(key, value):
    check(lib().rag_set_option(key.encode(), int(value)))


def set_stream(cuda_stream_ptr):
    check(lib().rag_set_stream(C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))
