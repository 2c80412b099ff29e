#!/usr/bin/env python
"""bench.py — the retrieval hot path of RAG-DPO on B200: queries/sec and latency of exact top-k @ 1024-d.

Headline workload (BASELINE.json configs[1], "C2"): 1M x 1024 fp32 synthetic corpus per GPU, one step = one batch
of B queries (default 1024) -> exact top-10.  At N > 1 GPUs (one process per GPU under torchrun) the corpus is
row-sharded (weak scaling: 1M rows per GPU), every rank scores the same batch and the local top-k lists are exchanged
over NVLink peer memory and merged.

`value`  queries/s with queries already resident in HBM (stream-ordered device call).
`e2e`    the same through the host-buffer C-ABI call: pinned H2D of the queries and D2H of ids/scores inside the
         timed region.
Units: at N=1 plain queries/s on the 1M-row corpus; at N>1 the corpus is N x 1M rows, so the job value is
queries/s x N ("1M-row-corpus equivalents"); the plain number is reported beside it as `queries_per_s`.

`configs` carries the other BASELINE.json configurations, each with its own time, roofline and in-run parity check
against the oracle (checker only):
  C1  50k x 1024 fp32, 48 questions x 4 query variants through the retriever API (N=1 only)
  C3  10M x 1024 bf16 row-sharded over the N GPUs, 4096-query batches, top-100 (strong scaling)
  C4  hybrid: 1M chunks, BM25 top-50 + dense top-50 per query, weighted RRF (N=1 only)
  C5  12.5M x 1024 bf16 rows PER GPU (N=8: the 100M-row north-star corpus), batch-1 latency and 4096-query batches

`--impl reference` times the CPU stand-in for the reference's own path (numpy fp32 brute force behind the
collection.query contract — chromadb itself is not installable here) on all host cores.
"""
import os
import sys


def _early_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs (reference arm, parity checks) must use the host cores.
    Has to happen before numpy loads its BLAS."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = os.cpu_count() or 1
    is_reference = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
    want = n if is_reference else max(1, n // world)     # rank 0 alone works in the reference arm
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(want)
    return want


CPU_THREADS_WANTED = _early_threads()

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

ROWS_PER_GPU = 1_000_000
DIM = 1024
TOPK = 10
CORPUS_SEED, QUERY_SEED = 1002, 2002
METRIC = "queries/sec & p50 latency, exact top-10 @ 1024-d, 1/2/4/8 B200; % HBM roofline"
UNIT = "queries/s (1M-row-corpus equivalents: corpus = n_gpus x 1M rows)"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled through NVML
    (nvidia_ml_py) every 20 ms; falls back to polling nvidia-smi when NVML is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            self.sm.append(float(c[1]))
            self.max_mhz = float(c[2])
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self._nvml else "nvidia-smi"}
        if self.power:
            out["power_w_max"] = float(max(self.power))
        return out


# ---------------------------------------------------------------------------
# CPU stand-in for the reference path (BASELINE.md §3 Ref-A) and the in-run parity checker
# ---------------------------------------------------------------------------
def cpu_topk(x, q, k):
    """numpy fp32 brute force: 1 - X @ q, argpartition + sort (the collection.query contract)."""
    s = q @ x.T                                           # (B, n) sgemm on all BLAS threads
    idx = np.argpartition(-s, k - 1, axis=1)[:, :k]
    part = np.take_along_axis(s, idx, axis=1)
    order = np.lexsort((idx, -part), axis=1)
    return np.take_along_axis(idx, order, axis=1), 1.0 - np.take_along_axis(part, order, axis=1)


def host_corpus(n, d, seed):
    g = np.random.default_rng(seed)
    x = np.empty((n, d), dtype=np.float32)
    for s in range(0, n, 65536):
        blk = g.standard_normal((min(65536, n - s), d), dtype=np.float32)
        blk /= np.linalg.norm(blk, axis=1, keepdims=True)
        x[s:s + len(blk)] = blk
    return x


def cpu_threads():
    """BLAS threads numpy actually uses (pinned to the wanted count: torchrun's OMP_NUM_THREADS=1 must not leak in)"""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=CPU_THREADS_WANTED)
        n = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
        return int(n)
    except Exception:
        return CPU_THREADS_WANTED


def time_cpu(x, q, k, batch, steps, warmup, budget_s):
    """returns (queries/s, ms per step, steps done)"""
    for _ in range(warmup):
        cpu_topk(x, q[:batch], k)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        cpu_topk(x, q[:batch], k)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return batch * done / dt, 1e3 * dt / done, done


def raw_rows(x32, dtype):
    """stored values (fp32) -> the storage-dtype bit patterns the C oracle reads"""
    from oracle import numpy_oracle as no
    if dtype == "f32":
        return np.ascontiguousarray(x32, np.float32), no.DT_F32
    if dtype == "bf16":
        return no.f32_to_bf16_bits(x32), no.DT_BF16
    return np.ascontiguousarray(x32).astype(np.float16).view(np.uint16), no.DT_F16


def oracle_topk_full(x32, q, k, dtype, pool=256):
    """CHECKER: exact top-k (canonical fp64 score desc, row asc) over ALL rows of x32 (the stored values widened
    to fp32).  An fp32 BLAS GEMM proposes the `pool` best rows per query, the C oracle (oracle/oracle.c) re-scores
    them in the canonical fp64 order; the pool provably contains the top-k when its worst fp32 score sits below
    the k-th exact score by more than the fp32 GEMM error (asserted)."""
    from oracle import c_oracle
    n = x32.shape[0]
    pool = min(pool, n)
    s = q @ x32.T
    idx = np.argpartition(-s, pool - 1, axis=1)[:, :pool] if pool < n else np.tile(np.arange(n), (len(q), 1))
    ids = np.empty((len(q), k), np.int64)
    scores = np.empty((len(q), k), np.float64)
    for b in range(len(q)):
        rows = np.sort(idx[b])
        raw, dt = raw_rows(x32[rows], dtype)
        sc = c_oracle.dense_scores(q[b], raw, dt)
        order = np.lexsort((rows, -sc))[:k]
        ids[b], scores[b] = rows[order], sc[order]
        if pool < n:
            worst = float(s[b, idx[b]].min())
            assert worst + 1e-4 < scores[b, -1], "fp32 candidate pool too shallow for the parity check"
    return ids, scores


def oracle_check_slices(corpus, q, rows, scores, dtype, slices, row_lo=0):
    """CHECKER for shards too large to brute-force on the host: (1) every returned (row, score) of the local
    result is re-scored by the C oracle from the downloaded row: bit-equal; (2) completeness on row slices: the
    rows of a slice whose oracle score reaches our k-th (score, row) must be exactly our returned rows inside the
    slice.  rows are LOCAL rows of `corpus`; returns the number of (query, row) pairs verified."""
    from oracle import c_oracle
    checked = 0
    uniq = np.unique(rows[rows >= 0])
    cache = {}
    for r in uniq.tolist():
        cache[r] = corpus.download(int(r), 1)[0]
    for b in range(len(q)):
        rr = rows[b][rows[b] >= 0]
        raw, dt = raw_rows(np.stack([cache[int(r)] for r in rr]), dtype)
        sc = c_oracle.dense_scores(q[b], raw, dt)
        assert np.array_equal(sc, scores[b, :len(rr)]), f"returned scores differ from the oracle (query {b})"
        assert all((scores[b, i] > scores[b, i + 1]) or (scores[b, i] == scores[b, i + 1] and rr[i] < rr[i + 1])
                   for i in range(len(rr) - 1)), f"result not ordered by (score desc, row asc) (query {b})"
        checked += len(rr)
    from concurrent.futures import ThreadPoolExecutor
    for lo, hi in slices:
        x = corpus.download(lo, hi - lo)
        raw, dt = raw_rows(x, dtype)

        def one(b):
            osc = c_oracle.dense_scores(q[b], raw, dt)           # ctypes releases the GIL: queries in parallel
            kth, kth_row = scores[b, -1], rows[b, -1]
            hit = np.nonzero(osc >= kth)[0]
            want = {int(lo + i): float(osc[i]) for i in hit if osc[i] > kth or lo + i <= kth_row}
            got = {int(r): float(s) for r, s in zip(rows[b], scores[b]) if lo <= r < hi}
            return got == want
        with ThreadPoolExecutor(max(1, CPU_THREADS_WANTED)) as ex:
            res = list(ex.map(one, range(len(q))))
        assert all(res), f"slice [{lo},{hi}): device rows of query {res.index(False)} differ from the oracle"
        checked += (hi - lo) * len(q)
    return checked


def workload_config(n_local, world, d, k, B, dtype):
    esz = 4 if dtype == "f32" else 2
    return {"workload": f"dense exact top-{k}: {n_local} x {d} {dtype} rows per GPU ({n_local * world} total), "
                        f"batch {B} queries per step",
            "corpus_rows": n_local * world, "rows_per_gpu": n_local, "dim": d, "k": k, "batch": B,
            "parallelism": f"row-shard x{world}, local top-k exchanged and merged" if world > 1 else "single GPU",
            "l2": f"corpus shard ({n_local * d * esz / 1e9:.1f} GB) is larger than L2 (126 MB): every step "
                  f"re-streams it from HBM"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from b200rag import synth
    n, d, k = ROWS_PER_GPU, DIM, TOPK
    cores = cpu_threads()
    x = host_corpus(n, d, CORPUS_SEED)
    sample_b = min(args.batch, 64)
    q = synth.unit_queries(sample_b, d, QUERY_SEED)
    qps, ms, done = time_cpu(x, q, k, sample_b, args.steps, max(1, min(args.warmup, 2)), budget_s=90.0)
    # the reference's real behaviour: it never batches, it issues B=1 calls one after the other
    # (src/rag/retriever.py:372-385)
    lat = []
    for i in range(min(12, sample_b)):
        t0 = time.perf_counter()
        cpu_topk(x, q[i:i + 1], k)
        lat.append(1e3 * (time.perf_counter() - t0))
    lat = lat[2:] if len(lat) > 4 else lat
    sample = (f"{done} steps of a {sample_b}-query numpy fp32 GEMM batch (X @ q, argpartition+sort) over the full "
              f"{n}x{d} fp32 corpus of ONE GPU's share on {cores} BLAS threads (os.cpu_count={os.cpu_count()}); the "
              f"value is per-query throughput on that share, i.e. already in 1M-row-corpus equivalents; chromadb "
              f"1.4.1 (HNSW) is not installable here, this is the exact search it approximates")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT,
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, world, d, k, args.batch, "f32"),
            "latency_b1_ms_p50": float(np.median(lat)),
            "sequential_b1": {"queries_per_s": 1e3 / float(np.mean(lat)), "calls": len(lat),
                              "note": "what the reference does: one collection.query call per query variant"},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
class Env:
    """process-wide state of the B200 arm: ranks, device, the one stream torch and the library share"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from b200rag import _lib
        self._lib = _lib
        _lib.lib()
        # one non-default stream for torch's ops AND the library's kernels, so that stream order is the only
        # synchronisation and torch.cuda.Event sees everything
        self.stream = torch.cuda.Stream(device=self.dev)
        _lib.set_stream(self.stream.cuda_stream)
        torch.cuda.set_stream(self.stream)
        self.peaks = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        t = self.torch.tensor([0.0 if ok else 1.0], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item()) == 0.0

    def timed(self, fn, steps, walls=None):
        """K calls of fn bracketed by barrier + synchronize on both sides, CUDA events on the shared stream,
        MAX over ranks.  Returns (total ms, kernel launches of the library inside the region)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = self._lib.counters()["launches"]
        e0.record(self.stream)
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            if walls is not None:
                walls.append(1e3 * (time.perf_counter() - t0))
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), self._lib.counters()["launches"] - c0


class DenseJob:
    """One dense configuration on this rank's shard: device-resident and host-buffer steps, parity check."""

    def __init__(self, env, n_local, dtype, seed, B, k, q_seed):
        from b200rag import _lib, synth
        from b200rag.sharded import ShardedDenseIndex
        self.env, self.n_local, self.dtype, self.B, self.k = env, n_local, dtype, B, k
        torch = env.torch
        self.index = ShardedDenseIndex(DIM, n_local * env.world, dtype=dtype, device=env.dev)
        self.index.fill_synthetic(seed)
        self.corpus = self.index.corpus
        self.q_host = _lib.pinned_empty((B, DIM), np.float32)
        self.q_host[:] = synth.unit_queries(B, DIM, q_seed)
        self.out_host = (_lib.pinned_empty((B, k), np.int32), _lib.pinned_empty((B, k), np.float64),
                         _lib.pinned_empty((B,), np.int32))
        self.q_dev = torch.from_numpy(np.array(self.q_host)).to(env.dev)
        self._steps = {}

    def device_step(self, nq=None):
        """whole hot path with HBM-resident queries: local top-k (+ exchange + merge when sharded); returns the
        function and the device tensors it fills: (ids, scores) merged over the ranks, and the local (rows, scores)"""
        nq = self.B if nq is None else nq
        if nq not in self._steps:
            self._steps[nq] = self.index.make_device_step(self.q_dev.data_ptr(), nq, self.k)
        return self._steps[nq]

    def local_step(self, nq=None):
        """this rank's local top-k only (no exchange, no merge): what the sharded step adds is read off the difference"""
        nq = self.B if nq is None else nq
        torch, k = self.env.torch, self.k
        key = ("local", nq)
        if key not in self._steps:
            o = (torch.empty((nq, k), dtype=torch.int32, device=self.env.dev),
                 torch.empty((nq, k), dtype=torch.float64, device=self.env.dev),
                 torch.empty((nq,), dtype=torch.int32, device=self.env.dev))
            qp = self.q_dev.data_ptr()

            def step():
                self.corpus.topk_dev(qp, nq, k, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
            self._steps[key] = (step, o)
        return self._steps[key][0]

    def exchange_cost(self, ms_step, steps=5):
        """sharded runs: per-rank time of the local top-k alone (min / max over ranks) and what exchange + merge (and
        waiting for the slowest rank) add to the step"""
        env = self.env
        if env.world == 1:
            return None
        fn = self.local_step()
        fn()
        torch = env.torch
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        for _ in range(steps):
            fn()
        e1.record(env.stream)
        env.barrier()
        mine = e0.elapsed_time(e1) / steps
        t = torch.tensor([mine, -mine], device=env.dev, dtype=torch.float64)
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
        hi, lo = float(t[0].item()), -float(t[1].item())
        return {"local_topk_ms_max_over_ranks": hi, "local_topk_ms_min_over_ranks": lo,
                "exchange_merge_ms": ms_step - hi}

    def host_step(self, nq=None):
        nq = self.B if nq is None else nq
        if self.env.world > 1:
            return self.index.topk(self.q_host[:nq], self.k)
        if nq == self.B:
            return self.corpus.topk(self.q_host, self.k, out=self.out_host)
        return self.corpus.topk(self.q_host[:nq], self.k)

    def exchange_kind(self):
        if self.env.world == 1:
            return None
        return ("peer-memory exchange (P2P stores over NVLink + epoch flags) + merge"
                if self.index.exchange(self.B, self.k) is not None else "NCCL all-gather + merge")

    def parity(self, nq, full):
        """compare the device result of the first nq queries with the oracle.  full: brute force over the whole
        shard on the host (fp32 GEMM pool + canonical fp64 re-score); else returned-row re-score + completeness on
        two row slices.  At N > 1 every rank checks its LOCAL top-k against the oracle over its own shard, the
        oracle's local lists are gathered and merged on the host, and rank 0 compares the device's merged
        result with that.  Returns a description; raises SystemExit on a mismatch."""
        env, k = self.env, self.k
        nq = min(nq, self.B)
        step, out = self.device_step(self.B)
        step()
        env.torch.cuda.synchronize()
        l_rows = out["local_rows"].cpu().numpy()[:nq]
        l_scores = out["local_scores"].cpu().numpy()[:nq]
        m_ids = out["ids"].cpu().numpy()[:nq].astype(np.int64)
        m_scores = out["scores"].cpu().numpy()[:nq]
        q = np.array(self.q_host[:nq])
        ok, what, err = True, "", ""
        try:
            if full:
                x = self.corpus.download()
                o_ids, o_scores = oracle_topk_full(x, q, k, self.dtype)
                del x
                assert np.array_equal(l_rows.astype(np.int64), o_ids), "local top-k ids differ from the oracle"
                assert np.array_equal(l_scores, o_scores), "local top-k scores differ from the oracle"
                what = (f"{nq} queries: ids and fp64 scores bit-equal to the oracle (fp32 GEMM pool of 256 rows per "
                        f"query re-scored in the canonical fp64 order by oracle.c) over all {self.n_local} rows")
            else:
                top = int(l_rows[0, 0]) // 100_000 * 100_000
                slices = [(top, min(self.n_local, top + 100_000)),
                          (self.n_local // 3, min(self.n_local, self.n_local // 3 + 50_000))]
                n_chk = oracle_check_slices(self.corpus, q, l_rows, l_scores, self.dtype, slices)
                o_ids, o_scores = l_rows.astype(np.int64), l_scores
                what = (f"{nq} queries: every returned (row, score) re-scored bit-equal by oracle.c, order checked, "
                        f"and completeness against the oracle on row slices {slices} ({n_chk} pairs)")
        except AssertionError as e:
            ok, err = False, str(e)
            o_ids = l_rows.astype(np.int64)
            o_scores = l_scores
        if env.world > 1:
            # merge the oracle's (verified) local lists on the host and compare with the device's merged result
            torch, dist = env.torch, env.dist
            g_ids = torch.empty((env.world, nq, k), dtype=torch.int64, device=env.dev)
            g_sc = torch.empty((env.world, nq, k), dtype=torch.float64, device=env.dev)
            gid = np.where(o_ids >= 0, o_ids + self.index.row_lo, -1)
            dist.all_gather_into_tensor(g_ids, torch.from_numpy(np.ascontiguousarray(gid)).to(env.dev))
            dist.all_gather_into_tensor(g_sc, torch.from_numpy(np.ascontiguousarray(o_scores)).to(env.dev))
            a_ids = g_ids.cpu().numpy().transpose(1, 0, 2).reshape(nq, -1)
            a_sc = g_sc.cpu().numpy().transpose(1, 0, 2).reshape(nq, -1)
            for b in range(nq):
                keep = a_ids[b] >= 0
                order = np.lexsort((a_ids[b][keep], -a_sc[b][keep]))[:k]
                if not (np.array_equal(a_ids[b][keep][order], m_ids[b, :len(order)]) and
                        np.array_equal(a_sc[b][keep][order], m_scores[b, :len(order)])):
                    ok, err = False, err or f"merged result of query {b} differs from the host merge of the oracle lists"
                    break
            what += f"; merged result over {env.world} ranks equal to the host merge of the oracle's local lists"
        if not env.all_ok(ok):
            if env.rank == 0 or not ok:
                log(f"PARITY MISMATCH on rank {env.rank}: {err}")
            raise SystemExit(f"parity check failed: {err or 'on another rank'}")
        return "ok: " + what

    def close(self):
        self.index.close()
        self.corpus.close()
        self._steps.clear()
        self.env.torch.cuda.empty_cache()


_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def profile_traffic(csv_name, kernel_substr):
    """`traffic` of a roofline object: dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, read
    from the committed `ncu --set full` raw page under profiles/ (the capture of the same workload shape; ncu cannot
    run inside the timed bench).  None when the file or the kernel is missing: never a pasted literal."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    try:
        rows = list(csv.reader(open(path)))
        h, u = rows[0], rows[1]
        kn, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        vals = [float(r[ir]) * _UNIT[u[ir]] + float(r[iw]) * _UNIT[u[iw]] for r in rows[2:] if kernel_substr in r[kn]]
        if not vals:
            return None
        return {"bytes_per_launch": float(np.mean(vals)), "launches_in_capture": len(vals),
                "source": f"profiles/{csv_name} ({kernel_substr})"}
    except (OSError, ValueError, KeyError, IndexError):
        return None


def tensor_roofline(env, B, n_local, main_ms, sustained=False):
    flops = 2.0 * B * n_local * DIM
    ach = flops / (main_ms / 1e3) / 1e12
    peak = env.peaks["bf16_tflops_sustained" if sustained else "bf16_tflops"]
    return {"bound": "tensor", "kernel": "dense_gemm_topk_kernel (tcgen05 bf16 contraction + fused top-k), main pass",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "frac_of_burst": ach / env.peaks["bf16_tflops"], "frac_of_sustained": ach / env.peaks["bf16_tflops_sustained"],
            "traffic": None,
            "peak_source": env.peaks["source"] + (" (sustained cuBLAS bf16: kernel timed inside a long loop)" if sustained
                                                  else " (burst cuBLAS bf16; kernel timed alone)"),
            "launches_per_step": 1, "avg_launch_ms": main_ms, "algorithmic_flops_per_launch": flops}


def hbm_roofline(env, kernel, bytes_per_launch, ms):
    ach = bytes_per_launch / (ms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": env.peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / env.peaks["hbm_gbs"], "traffic": None, "peak_source": env.peaks["source"],
            "avg_launch_ms": ms, "algorithmic_bytes_per_launch": bytes_per_launch}


def stage_medians(env, step, n):
    """library-internal CUDA-event stage times (same stream) over n calls: median main-pass / filter-stage ms"""
    main, stage, refine = [], [], []
    for _ in range(n):
        step()
        env.torch.cuda.synchronize()
        tm = env._lib.last_timings()
        main.append(float(tm[6]) if float(tm[6]) > 0 else float(tm[0]))
        stage.append(float(tm[0]))
        refine.append(float(tm[2]))
    return float(np.median(main)), float(np.median(stage)), float(np.median(refine))


def latency(env, fn, n, skip=3):
    torch = env.torch
    lat = []
    for _ in range(n + skip):
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        fn()
        e1.record(env.stream)
        torch.cuda.synchronize()
        lat.append(env.max_over_ranks(e0.elapsed_time(e1)))
    return lat[skip:]


def run_c2(env, args):
    """headline: BASELINE config 2 per GPU"""
    _lib = env._lib
    n_local, d, k, B = args.rows_per_gpu, DIM, args.k, args.batch
    esz = 4 if args.dtype == "f32" else 2
    job = DenseJob(env, n_local, args.dtype, CORPUS_SEED, B, k, QUERY_SEED)
    step, _ = job.device_step(B)
    step_b1, _ = job.device_step(1)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    walls_dev, walls_host = [], []
    with ClockSampler(env.local) as clocks:
        ms_total, launches = env.timed(step, args.steps, walls_dev)
        main_ms, stage_ms, refine_ms = stage_medians(env, step, min(args.steps, 10))
        for _ in range(2):
            job.host_step()
        ms_e2e, _ = env.timed(job.host_step, args.steps, walls_host)
    clock_summary = clocks.summary()
    # sustained: >= 2 s of back-to-back steps (power-capped steady state), with its own clock record
    n_sus = max(args.steps, int(args.sustained_s * 1e3 / max(ms_total / args.steps, 1e-3)))
    with ClockSampler(env.local) as clocks_sus:
        ms_sus, _ = env.timed(step, n_sus)
    sus_main, _, _ = stage_medians(env, step, 5)
    # batch-1 latency (the HBM-bound regime).  Default dispatch: an fp32 corpus is filtered through its bf16
    # shadow by the contraction kernel (half the bytes); the CUDA-core scan over the fp32 rows is timed too.
    lat = latency(env, step_b1, 30)
    b1_main, b1_stage, b1_refine = stage_medians(env, step_b1, 10)
    lat_host = []
    for _ in range(30):
        t0 = time.perf_counter()
        job.host_step(1)
        lat_host.append(1e3 * (time.perf_counter() - t0))
    _lib.set_option("tc_b1_shadow", 0)
    lat_scan = latency(env, step_b1, 20)
    _, scan_stage, _ = stage_medians(env, step_b1, 10)
    _lib.set_option("tc_b1_shadow", 1)

    parity = job.parity(64, full=n_local <= 2_000_000)
    ms_step = ms_total / args.steps
    qps = B / (ms_step / 1e3)
    qps_e2e = B / (ms_e2e / args.steps / 1e3)
    world = env.world
    tc_min = int(os.environ.get("B200RAG_TC_MIN_BATCH", "2"))
    if B >= tc_min:
        roof = tensor_roofline(env, B, n_local, main_ms)
        if (n_local, B, k, args.dtype) == (1_000_000, 1024, 10, "f32"):     # the shape the committed capture was taken on
            tr = profile_traffic("r2_dense_step_ncu_full_raw.csv", "dense_gemm_topk_kernel<1, 2>")
            if tr:
                roof["traffic"], roof["traffic_source"] = tr["bytes_per_launch"], tr["source"]
        roof["filter_stage_ms"] = stage_ms          # query prep, sample pass, threshold kernel, main pass
        roof["select_refine_ms"] = refine_ms
        roof["whole_step_frac"] = 2.0 * B * n_local * d / (ms_step / 1e3) / 1e12 / env.peaks["bf16_tflops"]
    else:
        roof = hbm_roofline(env, "dense_scan_kernel", n_local * d * esz, main_ms / ((B + 3) // 4))
    shadow_bytes = n_local * d * 2           # batch-1 default path: bf16 rows (the corpus itself or its shadow)
    line = {
        "metric": METRIC,
        "value": qps * world * (n_local / 1e6), "unit": UNIT,
        "queries_per_s": qps,
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (tcgen05 filter, fp32 accumulate) + f64 (exact refine of the candidates)",
        "storage_dtype": args.dtype, "data": "synthetic",
        "config": workload_config(n_local, world, d, k, B, args.dtype),
        "e2e": {"value": qps_e2e * world * (n_local / 1e6), "unit": UNIT, "queries_per_s": qps_e2e,
                "h2d_bytes_per_step": int(B * d * 4),
                # single GPU: int32 rows + f64 scores + counts; sharded: int64 global ids + f64 scores + counts
                "d2h_bytes_per_step": int(B * k * (16 if world > 1 else 12) + B * 4),
                "call_ms_p50": float(np.median(walls_host)), "call_ms_max": float(max(walls_host))},
        "step_call_ms": {"p50": float(np.median(walls_dev)), "max": float(max(walls_dev))},
        "gpu_launches": int(launches),
        "parity": parity,
        "exchange": job.exchange_kind(),
        "exchange_cost": job.exchange_cost(ms_step, min(args.steps, 10)),
        "sustained": {"seconds": ms_sus / 1e3, "steps": n_sus, "ms_per_step": ms_sus / n_sus,
                      "queries_per_s": B / (ms_sus / n_sus / 1e3), "main_pass_ms": sus_main,
                      "roofline_frac_of_sustained_peak": 2.0 * B * n_local * d / (sus_main / 1e3) / 1e12 /
                      env.peaks["bf16_tflops_sustained"],
                      "clocks": clocks_sus.summary()},
        "latency_b1": {"device_ms_p50": float(np.percentile(lat, 50)), "device_ms_p99": float(np.percentile(lat, 99)),
                       "host_call_ms_p50": float(np.percentile(lat_host, 50)),
                       "host_call_ms_p99": float(np.percentile(lat_host, 99)),
                       "stages_ms": {"filter": b1_stage, "refine": b1_refine},
                       "roofline": hbm_roofline(env, "dense_scan_kernel over the bf16 rows" if args.dtype == "bf16" else
                                                "dense_gemm_topk_kernel over the bf16 shadow (sample + main)",
                                                shadow_bytes, b1_stage),
                       "scan_path": {"device_ms_p50": float(np.percentile(lat_scan, 50)),
                                     "roofline": hbm_roofline(env, f"dense_scan_kernel over the {args.dtype} rows",
                                                              n_local * d * esz, scan_stage)}},
        "roofline": roof,
        "clocks": clock_summary,
    }
    if world == 1 and not args.no_cpu and n_local <= 2_000_000:
        x = job.corpus.download()
        sample_b = min(B, 64)
        cores = cpu_threads()
        cqps, cms, cdone = time_cpu(x, job.q_host, k, sample_b, 6, 1, budget_s=25.0)
        del x
        line["cpu_baseline"] = {"value": cqps, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"{cdone} batches of {sample_b} queries, numpy fp32 GEMM + argpartition over the "
                                          f"same {n_local}x{d} corpus (downloaded from the GPU), {cores} BLAS threads"}
    job.close()
    return line


def run_c3(env, args):
    """BASELINE config 3: 10M x 1024 bf16 row-sharded over the N GPUs, 4096-query batches, top-100 (strong scaling)"""
    total, B, k = args.c3_rows, 4096, 100
    n_local = (total + env.world - 1) // env.world
    job = DenseJob(env, n_local, "bf16", 1003, B, k, 2003)
    step, _ = job.device_step(B)
    for _ in range(2):
        step()
    with ClockSampler(env.local) as clocks:
        ms, launches = env.timed(step, 5)
        main_ms, stage_ms, refine_ms = stage_medians(env, step, 3)
        job.host_step()
        ms_e2e, _ = env.timed(job.host_step, 3)
    parity = job.parity(64, full=False)
    ms_step = ms / 5
    xc = job.exchange_cost(ms_step)
    out = {"workload": f"C3: {total} x {DIM} bf16 over {env.world} GPU(s) ({n_local} rows per GPU), batch {B}, top-{k}",
           "scaling": "strong", "ms": ms_step, "queries_per_s": B / (ms_step / 1e3),
           "e2e": {"queries_per_s": B / (ms_e2e / 3 / 1e3), "ms": ms_e2e / 3, "h2d_bytes_per_step": B * DIM * 4,
                   "d2h_bytes_per_step": B * k * 16 + B * 4},
           "roofline": dict(tensor_roofline(env, B, n_local, main_ms, sustained=True), filter_stage_ms=stage_ms,
                            select_refine_ms=refine_ms),
           "gpu_launches": int(launches), "parity": parity, "exchange": job.exchange_kind(), "exchange_cost": xc,
           "clocks": clocks.summary()}
    job.close()
    return out


def run_c5(env, args):
    """BASELINE config 5 shape: 12.5M x 1024 bf16 rows PER GPU (8 GPUs = the 100M-row corpus), top-10:
    batch-1 latency (HBM-bound) and 4096-query batches (tensor-bound)"""
    n_local, B, k = args.c5_rows_per_gpu, 4096, 10
    job = DenseJob(env, n_local, "bf16", 1005, B, k, 2005)
    step, _ = job.device_step(B)
    step_b1, _ = job.device_step(1)
    for _ in range(3):
        step_b1()
    with ClockSampler(env.local) as clocks:
        lat = latency(env, step_b1, 40)
        b1_main, b1_stage, b1_refine = stage_medians(env, step_b1, 10)
        lat_host = []
        for _ in range(20):
            t0 = time.perf_counter()
            job.host_step(1)
            lat_host.append(1e3 * (time.perf_counter() - t0))
        for _ in range(2):
            step()
        ms, launches = env.timed(step, 5)
        main_ms, stage_ms, refine_ms = stage_medians(env, step, 3)
        job.host_step()
        ms_e2e, _ = env.timed(job.host_step, 3)
    parity = job.parity(64, full=False)
    ms_step = ms / 5
    total = n_local * env.world
    out = {"workload": f"C5 shape: {n_local} x {DIM} bf16 rows per GPU x {env.world} GPU(s) = {total} rows "
                       f"({total * DIM * 2 / 1e9:.1f} GB), top-{k}" + (" — the north-star corpus" if env.world == 8 else
                                                                       " — one eighth-shard per GPU of the north-star corpus"),
           "scaling": "weak",
           "batch1": {"device_ms_p50": float(np.percentile(lat, 50)), "device_ms_p99": float(np.percentile(lat, 99)),
                      "host_call_ms_p50": float(np.percentile(lat_host, 50)),
                      "host_call_ms_p99": float(np.percentile(lat_host, 99)),
                      "target_ms": 5.58, "stages_ms": {"scan": b1_stage, "refine": b1_refine},
                      "roofline": hbm_roofline(env, "dense_scan_kernel (TMA bulk ring, fp32 FMA, fused top-k)",
                                               n_local * DIM * 2, b1_stage),
                      "aggregate_GBps": total * DIM * 2 / (float(np.percentile(lat, 50)) / 1e3) / 1e9},
           "batch4096": {"ms": ms_step, "queries_per_s": B / (ms_step / 1e3),
                         "e2e": {"queries_per_s": B / (ms_e2e / 3 / 1e3), "ms": ms_e2e / 3},
                         "roofline": dict(tensor_roofline(env, B, n_local, main_ms, sustained=True),
                                          filter_stage_ms=stage_ms, select_refine_ms=refine_ms),
                         "gpu_launches": int(launches)},
           "parity": parity, "exchange": job.exchange_kind(), "clocks": clocks.summary()}
    job.close()
    return out


def run_c1(env, args):
    """BASELINE config 1: 50k chunks x 1024 fp32, 48 questions x 4 query variants through the retriever API
    (HybridRetriever = mirror of RAGRetriever, src/rag/retriever.py:312-470), per question and batched, next to the
    same retriever logic around the CPU checkers (numpy fp32 collection + restated pure-Python rank-bm25)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from b200rag import DeviceCollection, DeviceChunkBM25Index, HybridRetriever, synth, tokenize_french

    class Provider:
        def __init__(self, table):
            self.table = table

        def embed(self, texts):
            return [self.table[t] for t in texts]

    class Expander:
        def expand(self, q):
            return [q, q + " reformulation une", q + " reformulation deux", q + " reformulation trois"]

    n, d, nq = args.c1_chunks, DIM, 48
    g = np.random.default_rng(1001)
    vocab = np.array([f"mot{i}" for i in range(30_000)])
    p = np.arange(1, len(vocab) + 1, dtype=np.float64) ** (-1.07)
    p /= p.sum()
    lens = g.integers(40, 251, size=n)
    flat = g.choice(len(vocab), size=int(lens.sum()), p=p)
    words = vocab[flat]
    ends = np.cumsum(lens)
    texts = [" ".join(words[e - ln:e]) for e, ln in zip(ends, lens)]
    emb = synth.synth_rows(1001, 0, n, d)
    metas = [{"document_path": f"doc_{i // 9}", "chunk_nature": "GUIDE", "chunk_index": i % 9, "confidence": "high",
              "source": "CNIL", "source_url": f"https://www.cnil.fr/fr/doc-{i // 9}"} for i in range(n)]
    ids = [f"doc{i // 9}_{i % 9}" for i in range(n)]
    questions = [" ".join(vocab[t] for t in g.choice(len(vocab), size=10, p=p)) + f" mot{2000 + i} mot{5000 + i}"
                 for i in range(nq)]
    qvec = synth.unit_queries(nq * 4, d, 2001)
    ex = Expander()
    table = {v: qvec[4 * i + j].tolist() for i, q in enumerate(questions) for j, v in enumerate(ex.expand(q))}
    t0 = time.perf_counter()
    col = DeviceCollection(dim=d, dtype="f32", capacity=n)
    for s in range(0, n, 5000):
        col.add(ids=ids[s:s + 5000], documents=texts[s:s + 5000], embeddings=emb[s:s + 5000], metadatas=metas[s:s + 5000])
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    bm = DeviceChunkBM25Index()
    bm.build_from_collection(col)
    t_build = time.perf_counter() - t0
    r = HybridRetriever(collection=col, embedding_provider=Provider(table), chunk_bm25_index=bm, query_expander=ex,
                        enable_summary_prefilter=False)
    r.retrieve_candidates(questions[0], n_candidates=40)
    lat, res_single = [], []
    for q in questions:
        t0 = time.perf_counter()
        res_single.append(r.retrieve_candidates(q, n_candidates=40))
        lat.append(1e3 * (time.perf_counter() - t0))
    r.retrieve_candidates_batch(questions, n_candidates=40)      # warm-up with the timed shape (lazy kernel loading, scratch)
    t_b = []
    for _ in range(3):
        t0 = time.perf_counter()
        res_batch = r.retrieve_candidates_batch(questions, n_candidates=40)
        t_b.append(time.perf_counter() - t0)
    t_batch = float(np.median(t_b))
    if os.environ.get("B200RAG_PROFILE_C1"):       # where the host time of the two front-ends goes (stderr)
        import cProfile
        import pstats
        for name, fn in (("per question", lambda: [r.retrieve_candidates(q, n_candidates=40) for q in questions]),
                         ("batch", lambda: r.retrieve_candidates_batch(questions, n_candidates=40))):
            pr = cProfile.Profile()
            pr.enable()
            fn()
            pr.disable()
            log(f"--- cProfile: {name}")
            pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(22)
    same = all([c.chunk_id for c in a] == [c.chunk_id for c in b] and
               [c.hybrid_score for c in a] == [c.hybrid_score for c in b] and
               [c.distance for c in a] == [c.distance for c in b] and
               [c.bm25_score for c in a] == [c.bm25_score for c in b] for a, b in zip(res_single, res_batch))
    # raw device calls of one question (what the reference issues: dense n_results=50, BM25 top_k=50)
    q4 = np.asarray([table[v] for v in ex.expand(questions[0])], dtype=np.float32)
    one = []
    for _ in range(30):
        t0 = time.perf_counter()
        col.query_rows(q4[:1], 50)
        one.append(1e3 * (time.perf_counter() - t0))
    dense_dev_ms = float(env._lib.last_timings()[0] + env._lib.last_timings()[2])
    toks = tokenize_french(questions[0])
    one_b = []
    for _ in range(30):
        t0 = time.perf_counter()
        bm.search_rows([toks], 50)
        one_b.append(1e3 * (time.perf_counter() - t0))
    # CPU checkers: same retriever logic around the exact numpy collection + restated pure-Python rank-bm25,
    # i.e. what the reference runs; its ids / scores must equal the device's
    import helpers
    from oracle import numpy_oracle as no
    xs = no.l2_normalize_rows(emb)

    class FastExact:
        """BASELINE.md Ref-A collection behind the collection.query contract: a numpy fp32 BLAS product proposes
        128 rows per query, oracle.c orders them by the canonical fp64 score (so near-ties resolve exactly)"""

        def count(self):
            return n

        def get(self, limit=None, offset=0, include=None, **kw):
            rows = range(offset, min(n, offset + limit) if limit is not None else n)
            return {"ids": [ids[i] for i in rows], "documents": [texts[i] for i in rows],
                    "metadatas": [metas[i] for i in rows]}

        def query(self, query_embeddings, n_results, where=None, include=None):
            q = no.l2_normalize_rows(np.asarray(query_embeddings, np.float32))
            o_ids, o_sc = oracle_topk_full(xs, q, n_results, "f32", pool=max(128, 2 * n_results))
            out = {"ids": [], "documents": [], "metadatas": [], "distances": []}
            for b in range(len(q)):
                idx = o_ids[b].tolist()
                out["ids"].append([ids[i] for i in idx]); out["documents"].append([texts[i] for i in idx])
                out["metadatas"].append([metas[i] for i in idx])
                out["distances"].append([no.distance_from_score(v) for v in o_sc[b]])
            return out

    cpu_lat, cpu_same, t_cpu_build = [], None, None
    if not args.no_cpu:
        t0 = time.perf_counter()
        obm = helpers.OracleChunkBM25Index(tokenize_french)
        obm.build_from_collection(FastExact())
        t_cpu_build = time.perf_counter() - t0
        rc = HybridRetriever(collection=FastExact(), embedding_provider=Provider(table), chunk_bm25_index=obm,
                             query_expander=ex, enable_summary_prefilter=False, fuse=helpers.oracle_fuse)
        cpu_same = True
        for qi, q in enumerate(questions[:args.c1_cpu_questions]):
            t0 = time.perf_counter()
            got = rc.retrieve_candidates(q, n_candidates=40)
            cpu_lat.append(1e3 * (time.perf_counter() - t0))
            cpu_same = cpu_same and [c.chunk_id for c in got] == [c.chunk_id for c in res_single[qi]] and \
                [c.bm25_score for c in got] == [c.bm25_score for c in res_single[qi]] and \
                [c.distance for c in got] == [c.distance for c in res_single[qi]] and \
                [c.hybrid_score for c in got] == [c.hybrid_score for c in res_single[qi]]
        if not cpu_same:
            raise SystemExit("C1 parity check failed: device retrieve_candidates differs from the CPU reference logic")
    if not same:
        raise SystemExit("C1: retrieve_candidates_batch differs from the per-question path")
    recall = None
    if not args.no_cpu and args.c1_recall_rows > 0:
        # recall@k of the reference's vector-store query against what the device returns (the exact top-k): the store
        # is the restated HNSW of oracle/hnsw.c with chromadb 1.4.1's defaults (M=16, ef_construction=100, ef_search=100,
        # cosine; a statistical twin, the wheel is absent), queried as the reference does (n_results=50, 192 variants)
        from b200rag import DeviceCorpus
        from oracle import c_oracle
        nr = min(args.c1_recall_rows, n)
        recall = {"store": "restated HNSW (oracle/hnsw.c): M=16, ef_construction=100, ef_search=100, cosine; PARITY UNPINNED",
                  "rows": nr, "queries": int(len(qvec)), "device_exact_recall": 1.0, "cases": []}
        gg = np.random.default_rng(7)
        docs_c = gg.standard_normal((nr // 10, d)).astype(np.float32)
        clustered = no.l2_normalize_rows(np.repeat(docs_c, 10, axis=0)[:nr] + 0.7 * gg.standard_normal((nr, d)).astype(np.float32))
        q_cl = no.l2_normalize_rows(docs_c[gg.choice(nr // 10, size=len(qvec), replace=True)] +
                                    0.7 * gg.standard_normal((len(qvec), d)).astype(np.float32))
        for name, xr, qr in (("config 1 synthetic (i.i.d. unit rows)", xs[:nr], qvec),
                             ("clustered twin (10 chunks per document)", clustered, q_cl)):
            dc = DeviceCorpus(d, "f32", capacity=nr)
            dc.append(np.ascontiguousarray(xr))
            rows_d, _, _ = dc.topk(np.ascontiguousarray(qr, dtype=np.float32), 50)
            dc.close()
            t0 = time.perf_counter()
            hx = c_oracle.HnswIndex(np.ascontiguousarray(xr), M=16, ef_construction=100)
            t_hb = time.perf_counter() - t0
            case = {"corpus": name, "hnsw_build_s": round(t_hb, 1)}
            for kk in (10, 50):
                t0 = time.perf_counter()
                ids_h, _ = hx.query(np.ascontiguousarray(qr, dtype=np.float32), kk, 100)
                case[f"hnsw_query_ms@{kk}"] = round(1e3 * (time.perf_counter() - t0) / len(qr), 3)
                case[f"recall@{kk}"] = round(float(np.mean([len(set(ids_h[i].tolist()) & set(rows_d[i, :kk].tolist())) / kk
                                                           for i in range(len(qr))])), 4)
            hx.close()
            recall["cases"].append(case)
    ms_q = float(np.percentile(lat, 50))
    return {"workload": f"C1: {n} chunks x {d} fp32, {nq} questions x 4 query variants (dense n_results=50 + BM25 top-50 "
                        f"per variant, weighted RRF, 40 candidates) through HybridRetriever.retrieve_candidates",
            "retrieve_candidates_ms_p50": ms_q, "retrieve_candidates_ms_p99": float(np.percentile(lat, 99)),
            "questions_per_s": 1e3 / ms_q,
            "retrieve_candidates_batch_ms_per_question": 1e3 * t_batch / nq,
            "batch_questions_per_s": nq / t_batch,
            "single_dense_query_call_ms_p50": float(np.percentile(one, 50)), "single_dense_query_device_ms": dense_dev_ms,
            "single_bm25_search_call_ms_p50": float(np.percentile(one_b, 50)),
            "device_load_s": t_load, "device_bm25_build_s": t_build,
            "roofline": {"bound": "hbm", "note": "latency regime: the 204.8 MB corpus streams in 31 us at the HBM peak; "
                         "a question is 8 device calls + host Python",
                         "achieved": n * d * 4 / (max(dense_dev_ms, 1e-6) / 1e3) / 1e9, "peak": env.peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": n * d * 4 / (max(dense_dev_ms, 1e-6) / 1e3) / 1e9 / env.peaks["hbm_gbs"]},
            "cpu_reference_logic_ms_per_question": float(np.median(cpu_lat)) if cpu_lat else None,
            "cpu_bm25_build_s": t_cpu_build,
            "recall_vs_reference_store": recall,
            "parity": (f"ok: batch == per-question path on {nq} questions (ids, distance, bm25 and hybrid scores); "
                       f"{len(cpu_lat)} questions equal to the reference retriever logic around the CPU checkers "
                       f"(ids, distance, bm25 and hybrid scores)") if cpu_lat else "ok: batch == per-question path"}


def zipf_tokens(n_docs, vocab, seed, lo=40, hi=250, s=1.07, threads=None):
    """synthetic tokenised corpus for the keyword leg: Zipf(s) term ids over `vocab` terms, doc length U[lo,hi];
    inverse-CDF sampling in parallel chunks (numpy releases the GIL).  Returns (doc_ptr int64 (n_docs+1),
    tokens int32, n_terms)."""
    from concurrent.futures import ThreadPoolExecutor
    g = np.random.default_rng(seed)
    lens = g.integers(lo, hi + 1, size=n_docs)
    doc_ptr = np.zeros(n_docs + 1, np.int64)
    np.cumsum(lens, out=doc_ptr[1:])
    total = int(doc_ptr[-1])
    p = np.arange(1, vocab + 1, dtype=np.float64) ** (-s)
    cdf = np.cumsum(p / p.sum())
    cdf[-1] = 1.0
    flat = np.empty(total, np.int32)
    threads = threads or max(1, min(32, (os.cpu_count() or 8)))
    bounds = np.linspace(0, total, threads * 4 + 1).astype(np.int64)

    def work(i):
        a, b = int(bounds[i]), int(bounds[i + 1])
        u = np.random.default_rng([seed, i]).random(b - a)
        flat[a:b] = np.searchsorted(cdf, u, side="right")
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(len(bounds) - 1)))
    np.minimum(flat, vocab - 1, out=flat)
    return doc_ptr, flat, vocab            # term id = Zipf rank - 1 (the synthetic vocabulary order)


def run_c4(env, args):
    """BASELINE config 4 (hybrid): 1M chunks, BM25 over CSR postings top-50 + dense top-50 per query variant,
    weighted RRF (k=60, the reference's weights) -> top-10 / top-40"""
    from b200rag import DeviceCorpus, rrf_fuse_rows, synth
    from b200rag.bm25 import DeviceBM25, Postings
    from oracle import c_oracle
    _lib = env._lib
    n_docs, vocab, Q = args.c4_docs, 200_000, 64
    t0 = time.perf_counter()
    doc_ptr, tokens, n_terms = zipf_tokens(n_docs, vocab, 1004)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    post = Postings.from_flat_tokens(doc_ptr, tokens, n_terms)
    t_host_build = time.perf_counter() - t0
    if not args.no_cpu:
        # the library's host CSR builder against the oracle's numpy restatement on the first 20k documents
        from oracle import numpy_oracle as no
        nd = min(20_000, n_docs)
        sub = Postings.from_flat_tokens(doc_ptr[:nd + 1], tokens[:doc_ptr[nd]], n_terms)
        o = no.CsrBM25(np.split(tokens[:doc_ptr[nd]].astype(np.int64), doc_ptr[1:nd]))
        v = len(o.term_ptr) - 1
        if not (np.array_equal(sub.term_ptr[:v + 1], o.term_ptr) and np.array_equal(sub.post_row, o.post_row) and
                np.array_equal(sub.post_tf, o.post_tf)):
            raise SystemExit("C4 parity check failed: rag_csr_build differs from the oracle's CSR")
    t0 = time.perf_counter()
    ix = DeviceBM25(post)
    t_dev_build = time.perf_counter() - t0
    g = np.random.default_rng(2004)
    p = np.arange(1, n_terms + 1, dtype=np.float64) ** (-1.07)
    p /= p.sum()
    # 4 query variants per question (original + 3 expansions), 8-12 Zipf terms + 2 mid-frequency terms
    queries = []
    for _ in range(Q * 4):
        qt = g.choice(n_terms, size=g.integers(8, 13), p=p)
        qt = np.concatenate([qt, g.integers(n_terms // 100, n_terms // 10, size=2)]).astype(np.int32)
        queries.append(qt)
    # algorithmic bytes of the filter pass: per token the term's list in the index format that serves it
    # (4 B per posting of the packed stream, or 2 B per row of the dense column of a frequent term)
    q_bytes, q_post = ix.query_bytes(queries)
    postings_per_query = float(q_post.mean())
    bytes_per_query = float(q_bytes.mean())
    bytes_per_posting = ix.bytes_per_posting()
    for q in queries[:3]:
        ix.search_ids([q], 50)
    lat, dev_ms = [], []
    for q in queries[:64]:
        t0 = time.perf_counter()
        ix.search_ids([q], 50)
        lat.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(float(_lib.last_timings()[0]))
    ix.search_ids(queries, 50)                      # warm-up (first large call allocates scratch)
    batch_wall, batch_dev = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        rows_b, scores_b, counts_b = ix.search_ids(queries, 50)
        batch_wall.append(time.perf_counter() - t0)
        batch_dev.append(float(_lib.last_timings()[0]))
    t_batch, dev_batch_ms = float(np.median(batch_wall)), float(np.median(batch_dev))
    # parity: oracle.c over the same CSR (fp64, numpy order), >= 16 queries, ids and scores bit-equal
    n_chk = 0 if args.no_cpu else 16
    cpu_ms = []
    for i in range(n_chk):
        t0 = time.perf_counter()
        want = c_oracle.bm25_scores(post.term_ptr, post.post_row, post.post_tf, post.doc_len, post.idf, post.avgdl,
                                    post.k1, post.b, queries[i])
        er, es = c_oracle.bm25_select(want, 50)
        cpu_ms.append(1e3 * (time.perf_counter() - t0))
        if not (rows_b[i, :counts_b[i]].tolist() == er.tolist() and np.array_equal(scores_b[i, :counts_b[i]], es)):
            raise SystemExit(f"C4 parity check failed: BM25 top-50 of query {i} differs from the oracle")
        r1, s1, c1 = ix.search_ids([queries[i]], 50)
        if not (r1[0, :c1[0]].tolist() == er.tolist() and np.array_equal(s1[0, :c1[0]], es)):
            raise SystemExit(f"C4 parity check failed: single-query BM25 path differs from the oracle (query {i})")
    # dense top-50 for the same questions (4 variants each) on a bf16 corpus of the same 1M chunks
    c = DeviceCorpus(DIM, "bf16", capacity=n_docs)
    c.fill_synthetic(seed=1004, nrows=n_docs)
    qv = synth.unit_queries(Q * 4, DIM, 2004)
    c.topk(qv[:8], 50)
    c.topk(qv, 50)                                  # warm-up (first call of this shape allocates scratch)
    dense_wall, dense_dev = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        rows_d, scores_d, counts_d = c.topk(qv, 50)
        dense_wall.append(time.perf_counter() - t0)
        tm = _lib.last_timings()
        dense_dev.append(float(tm[0]) + float(tm[2]))
    t_dense = float(np.median(dense_wall))
    lat_d4 = []
    for i in range(16):
        t0 = time.perf_counter()
        c.topk(qv[4 * i:4 * i + 4], 50)
        lat_d4.append(1e3 * (time.perf_counter() - t0))
    # RRF: rankings [dense q0, bm25 q0, dense q1, bm25 q1, ...], reference weights (retriever.py:374,405,431-432)
    ids = np.full((Q, 8, 50), -1, np.int32)
    for qi in range(Q):
        for v in range(4):
            ids[qi, 2 * v, :counts_d[4 * qi + v]] = rows_d[4 * qi + v, :counts_d[4 * qi + v]]
            ids[qi, 2 * v + 1, :counts_b[4 * qi + v]] = rows_b[4 * qi + v, :counts_b[4 * qi + v]]
    w = np.array([2.0, 3.0, 1.0, 0.75, 1.0, 0.75, 1.0, 0.75])
    rrf_fuse_rows(ids, w, 60, 10)
    t0 = time.perf_counter()
    fi, fs, fc = rrf_fuse_rows(ids, w, 60, 10)
    t_rrf = time.perf_counter() - t0
    for qi in range(min(Q, n_chk)):
        ei, es = c_oracle.rrf(ids[qi], w, 60, 10)
        if not (fi[qi, :fc[qi]].tolist() == ei.tolist() and np.array_equal(fs[qi, :fc[qi]], es)):
            raise SystemExit(f"C4 parity check failed: RRF of question {qi} differs from the oracle")
    ach = bytes_per_query * len(queries) / (dev_batch_ms / 1e3) / 1e9
    ach1 = bytes_per_query / (float(np.percentile(dev_ms, 50)) / 1e3) / 1e9
    # DRAM bytes of one 256-query filter launch from the committed capture of this shape (tools/prof_bm25.py)
    tr = profile_traffic("r2_bm25_filter_ncu_full_raw.csv", "bm25_filter") if (n_docs, len(queries)) == (1_000_000, 256) else None
    out = {"workload": f"C4 hybrid: {n_docs} chunks, vocab {n_terms}, doc length U[40,250], Zipf 1.07 ({len(post.post_row)} "
                       f"postings); {Q} questions x 4 query variants, BM25 top-50 + dense top-50 (bf16) each, RRF k=60",
           "bm25": {"batch_queries_per_s": len(queries) / t_batch, "batch_device_ms": dev_batch_ms,
                    "batch_device_us_per_query": 1e3 * dev_batch_ms / len(queries),
                    "single_query_call_ms_p50": float(np.percentile(lat, 50)),
                    "single_query_call_ms_p99": float(np.percentile(lat, 99)),
                    "single_query_device_ms_p50": float(np.percentile(dev_ms, 50)),
                    "avg_postings_per_query": postings_per_query, "bytes_per_posting": bytes_per_posting,
                    "roofline": {"bound": "hbm", "kernel": "bm25_filter_tma_kernel<16-bit accumulators>: integer filter pass over packed postings (TMA-staged runs) + dense columns (whole batched call: resolve + filter + finish)",
                                 "achieved": ach, "peak": env.peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": ach / env.peaks["hbm_gbs"],
                                 "traffic": tr["bytes_per_launch"] if tr else None,
                                 "traffic_source": tr["source"] if tr else None,
                                 "algorithmic_bytes_per_launch": bytes_per_query * len(queries),
                                 "algorithmic_bytes_per_query": bytes_per_query,
                                 "bytes_if_all_packed_4B_postings": postings_per_query * 4.0},
                    "roofline_single_query": {"bound": "hbm", "achieved": ach1, "peak": env.peaks["hbm_gbs"],
                                              "unit": "GB/s", "frac": ach1 / env.peaks["hbm_gbs"]},
                    "cpu_oracle_ms_per_query": float(np.median(cpu_ms)) if cpu_ms else None},
           "dense_top50": {"queries_per_s": len(qv) / t_dense, "batch_call_ms": 1e3 * t_dense,
                           "batch_device_ms_filter_plus_refine": float(np.median(dense_dev)), "four_variant_call_ms_p50": float(np.percentile(lat_d4, 50))},
           "rrf": {"questions_per_s": Q / t_rrf, "batch_ms": 1e3 * t_rrf},
           "hybrid_questions_per_s": Q / (t_batch + t_dense + t_rrf),
           "host_corpus_gen_s": t_gen, "host_csr_build_s": t_host_build, "device_build_s": t_dev_build,
           "parity": (f"ok: BM25 top-50 of {n_chk} queries (batched and single-query paths) and the fused top-10 of "
                      f"{min(Q, n_chk)} questions bit-equal to oracle.c over the same CSR") if n_chk else "skipped (--no-cpu)"}
    ix.close()
    c.close()
    env.torch.cuda.empty_cache()
    return out


def run_b200(args):
    env = Env()
    want = set(args.only.split(",")) if args.only else {"c2", "c1", "c3", "c4", "c5"}
    line = None
    configs = {}
    t_all = time.perf_counter()
    if "c2" in want:
        log("C2 headline ...")
        line = run_c2(env, args)
        log(f"C2 done in {time.perf_counter() - t_all:.1f} s")
    for name, fn, single_only in (("C3", run_c3, False), ("C5", run_c5, False), ("C1", run_c1, True), ("C4", run_c4, True)):
        if name.lower() not in want:
            continue
        if single_only and env.world > 1:
            configs[name] = {"skipped": "single-GPU configuration: measured by the N=1 run"}
            continue
        t0 = time.perf_counter()
        log(f"{name} ...")
        try:
            configs[name] = fn(env, args)
        except SystemExit:
            raise
        except Exception as e:                       # a failed side configuration must not lose the headline
            import traceback
            traceback.print_exc()
            configs[name] = {"error": repr(e)[:400]}
        if isinstance(configs[name], dict):
            configs[name]["leg_seconds"] = time.perf_counter() - t0
        log(f"{name} done in {time.perf_counter() - t0:.1f} s")
    env.barrier()
    if env.rank == 0:
        if line is None:
            line = {"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": env.world, "note": "headline skipped (--only)"}
        line["configs"] = configs
        line["bench_seconds"] = time.perf_counter() - t_all
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU,
                    help="rows per GPU shard of the headline workload (default: BASELINE config 2)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (cpu_baseline, CPU parity of C1/C4)")
    ap.add_argument("--only", default="", help="comma list of c2,c1,c3,c4,c5 (default: all)")
    ap.add_argument("--sustained-s", type=float, default=2.0)
    ap.add_argument("--c1-chunks", type=int, default=50_000)
    ap.add_argument("--c1-cpu-questions", type=int, default=3)
    ap.add_argument("--c1-recall-rows", type=int, default=10_000,
                    help="rows of the HNSW-twin recall case of C1 (0 = skip; its build is single-threaded CPU work)")
    ap.add_argument("--c3-rows", type=int, default=10_000_000)
    ap.add_argument("--c4-docs", type=int, default=1_000_000)
    ap.add_argument("--c5-rows-per-gpu", type=int, default=12_500_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
