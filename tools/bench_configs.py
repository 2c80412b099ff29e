#!/usr/bin/env python
"""Auxiliary measurements of the other BASELINE.json dense configs on ONE GPU (the per-GPU shard of the
multi-GPU configs).  Prints one JSON line per case.  Not the driver's bench (bench.py is)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {
    # name: (rows, dtype, batch, k, note)
    "c3_shard_g8": (1_250_000, "bf16", 4096, 100, "config 3 (10M x 1024 bf16, B=4096, top-100): one of 8 shards"),
    "c3_full_g1": (10_000_000, "bf16", 4096, 100, "config 3 on a single GPU"),
    "c5_shard_b1": (12_500_000, "bf16", 1, 10, "config 5 (100M x 1024 bf16 over 8 GPUs): one shard, batch-1 latency"),
    "c5_shard_b4096": (12_500_000, "bf16", 4096, 10, "config 5: one shard, 4096-query throughput"),
    "c2_b1": (1_000_000, "f32", 1, 10, "config 2 batch-1"),
    "c1_50k": (50_000, "f32", 1, 50, "config 1: 50k chunks, n_results=50 as the reference issues"),
    "c1_50k_b4": (50_000, "f32", 4, 50, "config 1: 4 query variants of one question, n_results=50"),
    "c1_17k_b4": (16_919, "f32", 4, 50, "the reference's real corpus size, 4 query variants, n_results=50"),
    "c2_b2": (1_000_000, "f32", 2, 10, "config 2 batch-2"),
    "c2_b4": (1_000_000, "f32", 4, 10, "config 2 batch-4"),
    "c2_b4_k50": (1_000_000, "f32", 4, 50, "config 2 batch-4 top-50"),
    "c2_b1_k50": (1_000_000, "f32", 1, 50, "config 2 batch-1 top-50"),
    "bf16_1m_b1": (1_000_000, "bf16", 1, 10, "1M bf16 batch-1"),
    "bf16_1m_b4_k50": (1_000_000, "bf16", 4, 50, "1M bf16 batch-4 top-50"),
    "c2_b8": (1_000_000, "f32", 8, 10, "config 2 batch-8"),
    "c2_b64": (1_000_000, "f32", 64, 10, "config 2 batch-64"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import torch
    from b200rag import DeviceCorpus, _lib, synth
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    d = 1024
    for name in args.cases:
        rows, dtype, B, k, note = CASES[name]
        c = DeviceCorpus(d, dtype, capacity=rows)
        c.fill_synthetic(seed=1000 + len(name), nrows=rows)
        q = synth.unit_queries(B, d, 2000 + len(name))
        dev = torch.device("cuda")
        qd = torch.from_numpy(q).to(dev)
        o_r = torch.empty((B, k), dtype=torch.int32, device=dev)
        o_s = torch.empty((B, k), dtype=torch.float64, device=dev)
        o_c = torch.empty((B,), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for _ in range(3):
            c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
        f0 = _lib.counters()["fallbacks"]
        lat, kern = [], []
        for _ in range(args.iters if B > 1 else 30):
            t0 = time.perf_counter()
            c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
            lat.append(1e3 * (time.perf_counter() - t0))
            kern.append([float(v) for v in _lib.last_timings()[:4]])
        lat_h = []
        for _ in range(args.iters if B > 1 else 30):
            t0 = time.perf_counter()
            c.topk(q, k)
            lat_h.append(1e3 * (time.perf_counter() - t0))
        main_ms = float(np.median([v[0] for v in kern]))
        esz = 4 if dtype == "f32" else 2
        out = {"case": name, "note": note, "rows": rows, "dtype": dtype, "batch": B, "k": k,
               "call_ms_p50": float(np.percentile(lat, 50)), "call_ms_p99": float(np.percentile(lat, 99)),
               "host_call_ms_p50": float(np.percentile(lat_h, 50)),
               "queries_per_s": B / (np.percentile(lat, 50) / 1e3),
               "stage_ms": {"main": main_ms, "merge": float(np.median([v[1] for v in kern])),
                            "refine": float(np.median([v[2] for v in kern]))},
               "fallback_calls": _lib.counters()["fallbacks"] - f0}
        if B >= 5:
            tf = 2.0 * B * rows * d / (main_ms / 1e3) / 1e12
            out["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                               "frac": tf / peaks["bf16_tflops"]}
        else:
            gb = rows * d * esz / (main_ms / 1e3) / 1e9
            out["roofline"] = {"bound": "hbm", "achieved": gb, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": gb / peaks["hbm_gbs"]}
        print(json.dumps(out), flush=True)
        c.close()
        del qd, o_r, o_s, o_c
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
