#!/usr/bin/env python
"""bench.py — queries/sec of the dense retrieval hot path (exact top-10 @ 1024-d).

Workload (BASELINE.json configs[1]): 1M x 1024 fp32 synthetic corpus per GPU,
one step = one batch of B queries (default 1024) -> exact top-10.  At N>1 GPUs
the corpus is row-sharded (weak scaling: 1M rows per GPU), every rank scores the
same batch, candidates are exchanged with one NCCL all-gather and merged.

`value`  queries/s with queries already resident in HBM (rag_dense_topk_dev).
`e2e`    the same through the host-buffer C-ABI call (rag_dense_topk): pinned
         H2D of the queries and D2H of ids/scores inside the timed region.
Units: at N=1 plain queries/s on the 1M-row corpus; at N>1 the corpus is N x 1M
rows, so the job value is queries/s x N ("1M-row-corpus equivalents") and the
plain number is reported beside it as `queries_per_s`.

`--impl reference` times the CPU stand-in for the reference's own path (numpy
fp32 brute force behind the collection.query contract — chromadb itself is not
installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

ROWS_PER_GPU = 1_000_000
DIM = 1024
TOPK = 10
CORPUS_SEED, QUERY_SEED = 1002, 2002
METRIC = "queries/sec & p50 latency, exact top-10 @ 1024-d, 1/2/4/8 B200; % HBM roofline"


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
# CPU stand-in for the reference path (BASELINE.md §3 Ref-A)
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
    try:
        from threadpoolctl import threadpool_info
        n = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


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


UNIT = "queries/s (1M-row-corpus equivalents: corpus = n_gpus x 1M rows)"
EXCHANGE_NOTE = {}      # how the ranks exchanged their candidates in this run (filled by run_b200)


def workload_config(n_local, world, d, k, B, dtype):
    esz = 4 if dtype == "f32" else 2
    return {"workload": f"dense exact top-{k}: {n_local} x {d} {dtype} rows per GPU ({n_local * world} total), "
                        f"batch {B} queries per step",
            "corpus_rows": n_local * world, "rows_per_gpu": n_local, "dim": d, "k": k, "batch": B,
            "parallelism": (f"row-shard x{world} + " + EXCHANGE_NOTE.get("kind", "all-gather merge")) if world > 1
            else "single GPU",
            "l2": f"corpus shard ({n_local * d * esz / 1e9:.1f} GB) is larger than L2 (126 MB): every step "
                  f"re-streams it from HBM"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from b200rag import synth
    n, d, k = ROWS_PER_GPU, DIM, TOPK
    x = host_corpus(n, d, CORPUS_SEED)
    sample_b = min(args.batch, 64)
    q = synth.unit_queries(sample_b, d, QUERY_SEED)
    qps, ms, done = time_cpu(x, q, k, sample_b, args.steps, max(1, min(args.warmup, 2)), budget_s=90.0)
    lat = []
    for i in range(min(8, sample_b)):
        t0 = time.perf_counter()
        cpu_topk(x, q[i:i + 1], k)
        lat.append(1e3 * (time.perf_counter() - t0))
    cores = cpu_threads()
    sample = (f"{done} steps of a {sample_b}-query numpy fp32 GEMM batch (X @ q, argpartition+sort) over the full "
              f"{n}x{d} fp32 corpus on {cores} BLAS threads (os.cpu_count={os.cpu_count()}); chromadb 1.4.1 (HNSW) is "
              f"not installable here, this is the exact search it approximates")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_config(n, world, d, k, args.batch, "f32")
    cfg["cpu_sample"] = (f"each step = one {sample_b}-query batch over a {n}-row corpus (one GPU's share of the workload); "
                         f"value is per-query throughput on that share, i.e. already in 1M-row-corpus equivalents")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT,
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "latency_b1_ms_p50": float(np.median(lat)),
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from b200rag import _lib, synth
    from b200rag.sharded import ShardedDenseIndex
    n_local, d, k, B = args.rows_per_gpu, DIM, args.k, args.batch
    esz = 4 if args.dtype == "f32" else 2
    n_total = n_local * world
    dev = torch.device("cuda", local)
    index = ShardedDenseIndex(d, n_total, dtype=args.dtype, device=dev)
    index.fill_synthetic(CORPUS_SEED)
    corpus = index.corpus
    # host-side query / result buffers live in page-locked memory (what a serving process would do)
    q_host = _lib.pinned_empty((B, d), np.float32)
    q_host[:] = synth.unit_queries(B, d, QUERY_SEED)
    out_host = (_lib.pinned_empty((B, k), np.int32), _lib.pinned_empty((B, k), np.float64),
                _lib.pinned_empty((B,), np.int32))
    q_dev = torch.from_numpy(np.array(q_host)).to(dev)
    L = _lib.lib()
    # one non-default stream for torch's collectives AND the library's kernels, so that
    # stream order is the only synchronisation and torch.cuda.Event sees everything
    stream = torch.cuda.Stream(device=dev)
    _lib.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)

    def make_device_step(nq):
        """whole hot path with HBM-resident queries: local top-k (+ exchange + merge when sharded)"""
        o_rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
        o_counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        # one packed buffer per rank: [scores (nq x k f64) | global ids (nq x k i64)] -> ONE all-gather
        mine = torch.empty((2, nq, k), dtype=torch.float64, device=dev)
        o_scores = mine[0]
        my_ids = mine[1].view(torch.int64)
        gathered = torch.empty((world, 2, nq, k), dtype=torch.float64, device=dev)
        m_scores = torch.empty((nq, k), dtype=torch.float64, device=dev)
        m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        m_counts = torch.empty((nq,), dtype=torch.int32, device=dev)

        ex = index.exchange(nq, k)       # peer-memory exchange (None: single GPU, or the NCCL all-gather path)

        def step():
            corpus.topk_dev(q_dev.data_ptr(), nq, k, o_rows.data_ptr(), o_scores.data_ptr(), o_counts.data_ptr())
            if world > 1:
                my_ids.copy_(o_rows)                       # int32 -> int64
                my_ids.add_(index.row_lo)                  # local row -> global id (k <= rows per shard: no padding)
                if ex is not None:                         # P2P stores into every peer's buffer + flags + merge
                    ex.merge_topk_dev(o_scores.data_ptr(), my_ids.data_ptr(), nq, k, m_scores.data_ptr(),
                                      m_ids.data_ptr(), m_counts.data_ptr())
                else:
                    dist.all_gather_into_tensor(gathered, mine)
                    _lib.check(L.rag_merge_topk_dev(gathered.data_ptr(), gathered.data_ptr() + nq * k * 8, world, nq,
                                                    k, 2 * nq * k, m_scores.data_ptr(), m_ids.data_ptr(),
                                                    m_counts.data_ptr()))
            return (m_ids, m_scores) if world > 1 else (o_rows, o_scores)
        return step

    step_device = make_device_step(B)
    step_device_b1 = make_device_step(1)
    if world > 1:
        EXCHANGE_NOTE["kind"] = ("peer-memory exchange (P2P stores over NVLink + epoch flags) + merge"
                                 if index.exchange(B, k) is not None else "NCCL all-gather + merge")

    def step_host():
        """the call a user makes: host buffers in, host results out"""
        if world > 1:
            return index.topk(q_host, k)
        return corpus.topk(q_host, k, out=out_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    call_ms = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.counters()["launches"]
        walls = []
        e0.record(torch.cuda.current_stream())
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            walls.append(1e3 * (time.perf_counter() - t0))
        e1.record(torch.cuda.current_stream())
        call_ms[fn.__name__] = walls
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.counters()["launches"] - c0

    for _ in range(max(args.warmup, 3)):
        step_device()
    with ClockSampler(local) as clocks:
        ms_total, launches = timed(step_device, args.steps)
        # kernel time of the dominant kernel, live (CUDA events inside the library, same stream)
        kern_ms, stage_ms = [], []
        for _ in range(min(args.steps, 10)):
            step_device()
            tm = _lib.last_timings()
            kern_ms.append(float(tm[6]) if float(tm[6]) > 0 else float(tm[0]))   # main pass alone (B >= 2)
            stage_ms.append(float(tm[0]))
        for _ in range(2):
            step_host()
        ms_e2e, _ = timed(step_host, args.steps)
        # batch-1 latency (the HBM-bound regime).  Default dispatch: an fp32 corpus is filtered through its bf16
        # shadow by the contraction kernel (half the bytes); the CUDA-core scan over the fp32 rows is timed too.
        def b1_latency():
            lat, stage = [], []
            for _ in range(30):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream())
                step_device_b1()
                e1.record(torch.cuda.current_stream())
                torch.cuda.synchronize()
                lat.append(e0.elapsed_time(e1))
                stage.append([float(v) for v in _lib.last_timings()[:4]])
            return lat[3:], stage[3:]
        lat, b1_stage_ms = b1_latency()
        lat_host = []
        for i in range(30):
            t0 = time.perf_counter()
            (index.topk(q_host[:1], k) if world > 1 else corpus.topk(q_host[:1], k))
            lat_host.append(1e3 * (time.perf_counter() - t0))
        _lib.set_option("tc_b1_shadow", 0)
        lat_scan, scan_stage_ms = b1_latency()
        _lib.set_option("tc_b1_shadow", 1)
    clock_summary = clocks.summary()

    ms_step = ms_total / args.steps
    qps = B / (ms_step / 1e3)
    qps_e2e = B / (ms_e2e / args.steps / 1e3)
    peaks = measured_peaks()
    # roofline of the dominant kernel of the step.  B >= 2: dense_gemm_topk_kernel (tcgen05 contraction +
    # fused select), tensor-bound: algorithmic flops = 2 * B * rows * dim per launch.  B <= 4: dense_scan_kernel,
    # HBM-bound: algorithmic bytes = rows * dim * sizeof(dtype) per launch (the shard is read once).
    tc_min = int(os.environ.get("B200RAG_TC_MIN_BATCH", "2"))
    main_ms = float(np.median(kern_ms))      # median: one NVML / driver hiccup must not skew the kernel figure
    bytes_per_launch = n_local * d * esz
    if B >= tc_min:
        flops = 2.0 * B * n_local * d
        ach = flops / (main_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "dense_gemm_topk_kernel (tcgen05 bf16 contraction + fused top-k)",
                "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "frac_of_sustained": ach / peaks["bf16_tflops_sustained"],
                # dram__bytes_read + write of the main pass, ncu --set full (profiles/r1_final_ncu_full_raw.csv)
                "traffic": 2.07e9 if (B == 1024 and n_local == 1_000_000 and args.dtype == "f32") else None,
                "peak_source": peaks["source"] + " (burst cuBLAS bf16; kernel timed alone)",
                "launches_per_step": 1, "avg_launch_ms": main_ms, "algorithmic_flops_per_launch": flops,
                # the whole filter stage around it: query prep, sample pass, threshold kernel, main pass
                "filter_stage_ms": float(np.median(stage_ms))}
    else:
        n_scan_launches = (B + 3) // 4
        scan_ms = main_ms / n_scan_launches
        ach = bytes_per_launch / (scan_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "dense_scan_kernel", "achieved": ach, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "launches_per_step": n_scan_launches, "avg_launch_ms": scan_ms,
                "algorithmic_bytes_per_launch": bytes_per_launch}
    b1_ms = float(np.median([v[0] for v in b1_stage_ms]))
    shadow_bytes = n_local * d * 2           # batch-1 default path: bf16 rows (the corpus itself or its shadow)
    achieved_b1 = shadow_bytes / (b1_ms / 1e3) / 1e9
    scan_ms = float(np.median([v[0] for v in scan_stage_ms]))
    achieved_scan = bytes_per_launch / (scan_ms / 1e3) / 1e9

    exchange_check = None
    if world > 1:
        # the peer-memory exchange must return what the NCCL all-gather + merge path returns (ids and fp64 scores)
        if index.exchange(B, k) is not None:
            got = index.topk(q_host[:16], k)
            os.environ["B200RAG_EXCHANGE"] = "nccl"
            want = index.topk(q_host[:16], k)
            os.environ["B200RAG_EXCHANGE"] = "peer"
            same = all(np.array_equal(a, b) for a, b in zip(got, want))
            exchange_check = "peer-memory exchange == NCCL all-gather path on 16 queries" if same else "MISMATCH"
            if not same:
                raise SystemExit("peer-memory exchange and NCCL path disagree")
        index.close()           # collective (barrier): every rank, before the non-zero ranks leave
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC,
        "value": qps * world * (n_local / 1e6), "unit": UNIT,
        "queries_per_s": qps,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (tcgen05 filter, fp32 accumulate) + f64 (exact refine of the candidates)",
        "storage_dtype": args.dtype, "data": "synthetic",
        "config": workload_config(n_local, world, d, k, B, args.dtype),
        "e2e": {"value": qps_e2e * world * (n_local / 1e6), "unit": UNIT,
                "queries_per_s": qps_e2e,
                "h2d_bytes_per_step": int(B * d * 4),
                # single GPU: int32 rows + f64 scores + counts; sharded: int64 global ids + f64 scores + counts
                "d2h_bytes_per_step": int(B * k * (16 if world > 1 else 12) + B * 4),
                "call_ms_p50": float(np.median(call_ms["step_host"])), "call_ms_max": float(max(call_ms["step_host"]))},
        "step_call_ms": {"p50": float(np.median(call_ms["step"])), "max": float(max(call_ms["step"]))},
        "gpu_launches": int(launches),
        "exchange_check": exchange_check,
        "latency_b1": {"device_ms_p50": float(np.percentile(lat, 50)), "device_ms_p99": float(np.percentile(lat, 99)),
                       "host_call_ms_p50": float(np.percentile(lat_host, 50)),
                       "host_call_ms_p99": float(np.percentile(lat_host, 99)),
                       "stages_ms": {"filter": b1_ms, "merge": float(np.median([v[1] for v in b1_stage_ms])),
                                     "refine": float(np.median([v[2] for v in b1_stage_ms]))},
                       "roofline": {"bound": "hbm",
                                    "kernel": ("dense_scan_kernel over the bf16 rows" if args.dtype == "bf16" else
                                               "dense_gemm_topk_kernel over the bf16 shadow (sample + main)"),
                                    "achieved": achieved_b1, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": achieved_b1 / peaks["hbm_gbs"], "traffic": None,
                                    "algorithmic_bytes_per_launch": shadow_bytes},
                       "scan_path": {"device_ms_p50": float(np.percentile(lat_scan, 50)), "scan_kernel_ms": scan_ms,
                                          "roofline": {"bound": "hbm", "kernel": f"dense_scan_kernel over the {args.dtype} rows",
                                                       "achieved": achieved_scan, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                       "frac": achieved_scan / peaks["hbm_gbs"],
                                                       "traffic": 4.096e9 if (args.dtype == "f32" and n_local == 1_000_000) else None,
                                                       "algorithmic_bytes_per_launch": bytes_per_launch}}},
        "roofline": roof,
        "clocks": clock_summary,
    }
    if world == 1 and not args.no_cpu and n_local <= 2_000_000:
        x = corpus.download()
        sample_b = min(B, 64)
        cqps, cms, cdone = time_cpu(x, q_host, k, sample_b, 6, 1, budget_s=25.0)
        # spot-check the CPU leg against the device result while we are here (ids only; fp32 BLAS scores)
        cores = cpu_threads()
        line["cpu_baseline"] = {"value": cqps, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"{cdone} batches of {sample_b} queries, numpy fp32 GEMM + argpartition over the "
                                          f"same {n_local}x{d} corpus (downloaded from the GPU), {cores} BLAS threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU,
                    help="rows per GPU shard (default: BASELINE config 2; 12500000 with --dtype bf16 = config 5)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
