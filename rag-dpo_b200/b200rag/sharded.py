"""Row-sharded dense search: one process per GPU, candidates merged with one
all-gather (NCCL over NVLink on the B200 box; gloo in the CPU tests).

The reference is single-process; this is the B200 scale-out of
collection.query (src/rag/retriever.py:215-220, 380-385): rank g owns the
contiguous rows [g*ceil(N/G), (g+1)*ceil(N/G)), every rank scores the same
query batch against its shard, and the exact (score, global id) top-k lists
are exchanged and merged with ties -> lowest global id.  Exactness: the global
top-k is a subset of the union of the local top-k lists.
"""
import numpy as np


def shard_bounds(n_rows, world, rank):
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def merge_lists_torch(scores, ids, k):
    """(G,B,k) -> (B,k) by (score desc, id asc); ids < 0 are padding.  Host-side
    restatement used only where no GPU exists (gloo tests)."""
    import torch
    G, B, kk = scores.shape
    s = scores.permute(1, 0, 2).reshape(B, G * kk)
    i = ids.permute(1, 0, 2).reshape(B, G * kk)
    s = torch.where(i >= 0, s, torch.full_like(s, float("-inf")))
    big = torch.where(i >= 0, i, torch.full_like(i, 2 ** 62))
    order1 = torch.argsort(big, dim=1, stable=True)
    s1, i1 = torch.gather(s, 1, order1), torch.gather(big, 1, order1)
    order2 = torch.argsort(s1, dim=1, descending=True, stable=True)
    s2, i2 = torch.gather(s1, 1, order2)[:, :k], torch.gather(i1, 1, order2)[:, :k]
    valid = i2 < 2 ** 62
    return torch.where(valid, s2, torch.zeros_like(s2)), torch.where(valid, i2, torch.full_like(i2, -1)), valid.sum(1).to(torch.int32)


class PeerExchange:
    """The exchange step over NVLink peer memory (csrc/exchange.cu): every rank stores its local top-k block into
    every peer's buffer and the merge kernel waits for the epoch flags — no collective library call per step.
    One instance per process; `slot_bytes` bounds B*k*16 of a call."""

    def __init__(self, world, rank, slot_bytes, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        self._lib, self._L = _lib, _lib.lib()
        self.world, self.rank, self.slot_bytes, self.group = world, rank, int(slot_bytes), group
        self._h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _lib.check(self._L.rag_exchange_create(C.byref(self._h), world, rank, self.slot_bytes, handle))
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(handle), group=group)
        blob = b"".join(gathered)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _lib.check(self._L.rag_exchange_connect(self._h, buf))
        dist.barrier(group=group)                      # every rank has mapped every buffer

    def merge_topk_dev(self, my_scores_ptr, my_ids_ptr, B, k, out_scores_ptr, out_ids_ptr, out_counts_ptr):
        self._lib.check(self._L.rag_exchange_merge_topk_dev(self._h, my_scores_ptr, my_ids_ptr, int(B), int(k),
                                                            out_scores_ptr, out_ids_ptr, out_counts_ptr))

    def merge_rows_dev(self, my_scores_ptr, my_rows_ptr, row_lo, B, k, out_scores_ptr, out_ids_ptr, out_counts_ptr,
                       my_counts_ptr=None):
        """same, the rank hands over its LOCAL int32 rows (the push kernel turns them into global ids on the way)
        and its counts: a query some rank left unresolved (-1) comes out with count -1 on every rank"""
        self._lib.check(self._L.rag_exchange_merge_rows_dev(self._h, my_scores_ptr, my_rows_ptr, int(row_lo),
                                                            my_counts_ptr, int(B), int(k), out_scores_ptr, out_ids_ptr,
                                                            out_counts_ptr))

    def close(self):
        if self._h:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)         # nobody is still pushing into a buffer that goes away
            self._L.rag_exchange_destroy(self._h)
            self._h = None


class ShardedDenseIndex:
    """local_topk(q32 np (B,d), k) -> (rows int32 (B,k) [-1 pad], scores f64 (B,k), counts).
    On the GPU box leave local_topk/merge at None: the shard lives in a
    DeviceCorpus and the merge runs in rag_merge_topk_dev."""

    def __init__(self, dim, n_rows_total, dtype="bf16", group=None, local_topk=None, merge=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dim = dim
        self.n_rows_total = int(n_rows_total)
        self.row_lo, self.row_hi = shard_bounds(self.n_rows_total, self.world, self.rank)
        self._local_topk = local_topk
        self._merge = merge
        self.corpus = None
        if local_topk is None:
            from . import _lib
            from .collection import DeviceCorpus
            self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            self.corpus = DeviceCorpus(dim, dtype, capacity=self.row_hi - self.row_lo)
        else:
            self.device = torch.device("cpu")
        self._exchange = None                    # PeerExchange, created on first use (GPU path, world > 1)
        self._exchange_failed = False

    def exchange(self, B, k):
        """the peer-memory exchange for calls up to this size, or None (single rank, B200RAG_EXCHANGE=nccl, or the
        buffers could not be shared between the processes: the NCCL all-gather + merge path is used instead)"""
        import os
        if self.world == 1 or self.corpus is None or self._exchange_failed:
            return None
        if os.environ.get("B200RAG_EXCHANGE", "peer") == "nccl":
            return None
        need = B * k * 16
        if self._exchange is not None and self._exchange.slot_bytes >= need:
            return self._exchange
        try:
            if self._exchange is not None:
                self._exchange.close()
                self._exchange = None
            self._exchange = PeerExchange(self.world, self.rank, max(need, 1 << 20), group=self.group)
        except Exception as exc:                      # every rank fails alike (same box, same driver)
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({exc}); using the NCCL all-gather")
            self._exchange_failed = True
            return None
        return self._exchange

    def close(self):
        if self._exchange is not None:
            self._exchange.close()
            self._exchange = None

    def fill_synthetic(self, seed):
        """each rank generates exactly its own rows of the global synthetic corpus"""
        self.corpus.fill_synthetic(seed, self.row_hi - self.row_lo, gen_row0=self.row_lo)

    def topk(self, q32, k):
        """q32: numpy (B,dim) fp32, identical on every rank.  Returns numpy
        (ids int64 (B,k) [-1 pad], scores f64 (B,k), counts int32 (B,))."""
        if self._local_topk is None and self.corpus is not None:
            return self._topk_device(q32, int(k))
        return self._topk_host(q32, int(k))

    def _topk_device(self, q32, kk, exact_local=False):
        """GPU path: queries H2D once, local top-k with device-resident outputs (stream-ordered call), the
        exchange over NVLink peer memory (or ONE packed NCCL all-gather) + merge, one D2H of the merged result.
        Queries some rank could not finish exactly inside the stream-ordered call (more deep ties than its
        device-driven fallback serves; the exchange marks them -1 on every rank) are redone by all ranks together
        with the host-checked local call (exact_local)."""
        torch, dist = self.torch, self.dist
        from . import _lib
        L = _lib.lib()
        B = q32.shape[0]
        dev = self.device
        n_local = self.row_hi - self.row_lo
        kl = min(kk, n_local)
        stream = torch.cuda.current_stream(dev)
        if stream.cuda_stream != 0:
            _lib.set_stream(stream.cuda_stream)       # library kernels and NCCL on the same stream
        shared = stream.cuda_stream != 0              # library kernels and NCCL run on this very stream
        buf = self._buffers(B, kk, kl)
        qt = torch.from_numpy(np.ascontiguousarray(q32, dtype=np.float32))
        if qt.is_pinned():                            # page-locked caller memory (rag_host_alloc): DMA it directly
            buf["q_dev"].copy_(qt, non_blocking=True)
        else:
            buf["q_host"].copy_(qt)
            buf["q_dev"].copy_(buf["q_host"], non_blocking=True)
        mine, my_ids = buf["mine"], buf["my_ids"]
        if not shared:
            stream.synchronize()                      # inputs are in place before the library's own stream reads them
        if exact_local:
            r, s_, c = self.corpus.topk(np.ascontiguousarray(q32, dtype=np.float32), kl) if kl > 0 else (None, None, None)
            mine[0].zero_()
            my_ids.fill_(-1)
            if kl > 0:
                mine[0, :, :kl] = torch.from_numpy(s_).to(dev)
                gid = torch.from_numpy(r.astype(np.int64)).to(dev)
                my_ids[:, :kl] = torch.where(gid >= 0, gid + self.row_lo, gid)
        elif kl == kk:
            # the local scores land directly in the packed exchange buffer; k <= rows per shard: no padding
            self.corpus.topk_dev(buf["q_dev"].data_ptr(), B, kl, buf["o_rows"].data_ptr(), mine[0].data_ptr(),
                                 buf["o_counts"].data_ptr())
            if not shared:
                _lib.sync_stream_of(torch, dev)       # the call is stream-ordered on the library's own stream
            if not (shared and self.exchange(B, kk) is not None):     # the peer exchange converts rows on the way out
                my_ids.copy_(buf["o_rows"])           # int32 -> int64
                my_ids.add_(self.row_lo)              # local row -> global id
        else:
            mine[0].zero_()
            my_ids.fill_(-1)
            if kl > 0:
                self.corpus.topk_dev(buf["q_dev"].data_ptr(), B, kl, buf["o_rows"].data_ptr(),
                                     buf["o_scores"].data_ptr(), buf["o_counts"].data_ptr())
                if not shared:
                    _lib.sync_stream_of(torch, dev)
                mine[0, :, :kl] = buf["o_scores"]
                gid = buf["o_rows"].to(torch.int64)
                my_ids[:, :kl] = torch.where(gid >= 0, gid + self.row_lo, gid)
        ex = self.exchange(B, kk) if shared else None
        unresolved_local = None
        if not exact_local and kl > 0 and not (ex is not None and kl == kk):
            # (the peer exchange carries the flag in-band: no extra kernels on the step's critical path)
            unresolved_local = (buf["o_counts"] < 0).any().to(torch.int32).reshape(1)
        if ex is not None:
            # stores into the peers' buffers + epoch flags + merge: two launches, no collective call
            if not exact_local and kl == kk:
                ex.merge_rows_dev(mine[0].data_ptr(), buf["o_rows"].data_ptr(), self.row_lo, B, kk,
                                  buf["m_scores"].data_ptr(), buf["m_ids"].data_ptr(), buf["m_counts"].data_ptr(),
                                  buf["o_counts"].data_ptr())
            else:
                ex.merge_topk_dev(mine[0].data_ptr(), my_ids.data_ptr(), B, kk, buf["m_scores"].data_ptr(),
                                  buf["m_ids"].data_ptr(), buf["m_counts"].data_ptr())
        else:
            if self.world > 1:
                gathered = buf["gathered"]
                dist.all_gather_into_tensor(gathered, mine, group=self.group)
            else:
                gathered = mine[None]
            if not shared:
                stream.synchronize()                  # the gather has landed before the library's stream merges
            _lib.check(L.rag_merge_topk_dev(gathered.data_ptr(), gathered.data_ptr() + B * kk * 8, self.world, B, kk,
                                            2 * B * kk, buf["m_scores"].data_ptr(), buf["m_ids"].data_ptr(),
                                            buf["m_counts"].data_ptr()))
        h = buf["host_out"]
        if not shared:
            _lib.sync_stream_of(torch, dev)
        h[0].copy_(buf["m_ids"], non_blocking=True)
        h[1].copy_(buf["m_scores"], non_blocking=True)
        h[2].copy_(buf["m_counts"], non_blocking=True)
        redo_all = False
        if unresolved_local is not None:
            # paths without the in-band flag (NCCL all-gather, padded lists): agree with one tiny all-reduce
            if self.world > 1:
                dist.all_reduce(unresolved_local, op=dist.ReduceOp.MAX, group=self.group)
            redo_all = bool(unresolved_local.item())
        stream.synchronize()
        ids, scores, counts = h[0].numpy().copy(), h[1].numpy().copy(), h[2].numpy().copy()
        bad = np.arange(B) if redo_all else np.nonzero(counts < 0)[0]
        if len(bad):
            r_ids, r_scores, r_counts = self._topk_device(np.ascontiguousarray(q32[bad]), kk, exact_local=True)
            ids[bad], scores[bad], counts[bad] = r_ids, r_scores, r_counts
        return ids, scores, counts

    def make_device_step(self, q_dev_ptr, nq, k):
        """Stream-ordered device step for HBM-resident queries (bench / serving loops): local top-k, then — when
        sharded — the exchange of the (score, global id) lists over NVLink peer memory and the merge.  Returns
        (step, out): step() only queues work on torch's current stream (which must be the library's stream,
        _lib.set_stream); out holds the device tensors it fills: "ids"/"scores"/"counts" (merged over the ranks;
        the local rows on a single rank) and "local_rows"/"local_scores"."""
        torch, dist = self.torch, self.dist
        from . import _lib
        L = _lib.lib()
        dev, world = self.device, self.world
        if k > self.row_hi - self.row_lo:
            raise ValueError("make_device_step needs k <= rows per shard")
        o_rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
        o_counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        # one packed buffer per rank: [scores (nq x k f64) | global ids (nq x k i64)] -> ONE all-gather on the NCCL path
        mine = torch.empty((2, nq, k), dtype=torch.float64, device=dev)
        o_scores = mine[0]
        my_ids = mine[1].view(torch.int64)
        out = {"ids": o_rows, "scores": o_scores, "counts": o_counts, "local_rows": o_rows, "local_scores": o_scores}
        if world == 1:
            def step():
                self.corpus.topk_dev(q_dev_ptr, nq, k, o_rows.data_ptr(), o_scores.data_ptr(), o_counts.data_ptr())
            return step, out
        m_scores = torch.empty((nq, k), dtype=torch.float64, device=dev)
        m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        m_counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        out.update(ids=m_ids, scores=m_scores, counts=m_counts)
        ex = self.exchange(nq, k)       # peer-memory exchange, or None: the NCCL all-gather path
        if ex is not None:
            def step():
                self.corpus.topk_dev(q_dev_ptr, nq, k, o_rows.data_ptr(), o_scores.data_ptr(), o_counts.data_ptr())
                # P2P stores of (scores | local rows -> global ids) into every peer's buffer + flags + merge
                ex.merge_rows_dev(o_scores.data_ptr(), o_rows.data_ptr(), self.row_lo, nq, k, m_scores.data_ptr(),
                                  m_ids.data_ptr(), m_counts.data_ptr(), o_counts.data_ptr())
            return step, out
        gathered = torch.empty((world, 2, nq, k), dtype=torch.float64, device=dev)

        def step():
            self.corpus.topk_dev(q_dev_ptr, nq, k, o_rows.data_ptr(), o_scores.data_ptr(), o_counts.data_ptr())
            my_ids.copy_(o_rows)                       # int32 -> int64
            my_ids.add_(self.row_lo)                   # local row -> global id (k <= rows per shard: no padding)
            dist.all_gather_into_tensor(gathered, mine, group=self.group)
            _lib.check(L.rag_merge_topk_dev(gathered.data_ptr(), gathered.data_ptr() + nq * k * 8, world, nq, k,
                                            2 * nq * k, m_scores.data_ptr(), m_ids.data_ptr(), m_counts.data_ptr()))
        return step, out

    def _buffers(self, B, kk, kl):
        """device / pinned buffers reused across calls of the same shape"""
        torch = self.torch
        key = (B, kk, kl)
        cache = self.__dict__.setdefault("_buf_cache", {})
        buf = cache.get(key)
        if buf is None:
            dev = self.device
            mine = torch.zeros((2, B, kk), dtype=torch.float64, device=dev)
            my_ids = mine[1].view(torch.int64)
            my_ids.fill_(-1)
            buf = {
                "q_host": torch.empty((B, self.dim), dtype=torch.float32).pin_memory(),
                "q_dev": torch.empty((B, self.dim), dtype=torch.float32, device=dev),
                "mine": mine, "my_ids": my_ids,
                "o_rows": torch.empty((B, max(kl, 1)), dtype=torch.int32, device=dev),
                "o_scores": torch.empty((B, max(kl, 1)), dtype=torch.float64, device=dev),
                "o_counts": torch.empty((B,), dtype=torch.int32, device=dev),
                "gathered": torch.empty((self.world, 2, B, kk), dtype=torch.float64, device=dev),
                "m_scores": torch.empty((B, kk), dtype=torch.float64, device=dev),
                "m_ids": torch.empty((B, kk), dtype=torch.int64, device=dev),
                "m_counts": torch.empty((B,), dtype=torch.int32, device=dev),
                "host_out": (torch.empty((B, kk), dtype=torch.int64).pin_memory(),
                             torch.empty((B, kk), dtype=torch.float64).pin_memory(),
                             torch.empty((B,), dtype=torch.int32).pin_memory()),
            }
            if len(cache) > 8:
                cache.clear()
            cache[key] = buf
        return buf

    def _topk_host(self, q32, kk):
        torch, dist = self.torch, self.dist
        B = q32.shape[0]
        if self._local_topk is not None:
            rows, scores, counts = self._local_topk(q32, kk)
        else:
            n_local = self.row_hi - self.row_lo
            kl = min(kk, n_local)
            rows = np.full((B, kk), -1, np.int32)
            scores = np.zeros((B, kk), np.float64)
            if kl > 0:
                r, s, c = self.corpus.topk(q32, kl)
                rows[:, :kl], scores[:, :kl] = r, s
        ids = np.where(rows >= 0, rows.astype(np.int64) + self.row_lo, -1)
        t_scores = torch.from_numpy(np.ascontiguousarray(scores)).to(self.device)
        t_ids = torch.from_numpy(np.ascontiguousarray(ids)).to(self.device)
        if self.world > 1:
            # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
            g_scores = torch.empty((self.world * B, kk), dtype=torch.float64, device=self.device)
            g_ids = torch.empty((self.world * B, kk), dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(g_scores, t_scores, group=self.group)
            dist.all_gather_into_tensor(g_ids, t_ids, group=self.group)
            g_scores = g_scores.view(self.world, B, kk)
            g_ids = g_ids.view(self.world, B, kk)
        else:
            g_scores, g_ids = t_scores[None], t_ids[None]
        if self._merge is not None or self.corpus is None:
            merge = self._merge or merge_lists_torch
            o_s, o_i, o_c = merge(g_scores, g_ids, kk)
        else:
            from . import _lib
            torch.cuda.current_stream(self.device).synchronize()     # the gather must have landed
            o_s = torch.empty((B, kk), dtype=torch.float64, device=self.device)
            o_i = torch.empty((B, kk), dtype=torch.int64, device=self.device)
            o_c = torch.empty((B,), dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib().rag_merge_topk_dev(g_scores.data_ptr(), g_ids.data_ptr(), self.world, B, kk, 0,
                                                     o_s.data_ptr(), o_i.data_ptr(), o_c.data_ptr()))
            _lib.sync_stream_of(torch, self.device)
        return o_i.cpu().numpy(), o_s.cpu().numpy(), o_c.cpu().numpy()
