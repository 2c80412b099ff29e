#!/usr/bin/env python
"""Cost of the NVML queries bench.py's clock sampler makes, and their effect on the mean step time."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    sys.path.insert(0, p)
import torch, pynvml
from b200rag import DeviceCorpus, _lib, synth
import bench
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def cost(f, n=20):
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); t.append(1e3 * (time.perf_counter() - t0))
    return [round(float(np.median(t)), 3), round(float(max(t)), 3)]
print(json.dumps({"clock_ms": cost(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                  "power_ms": cost(lambda: pynvml.nvmlDeviceGetPowerUsage(h)),
                  "reasons_ms": cost(lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))}), flush=True)
rows, d, B, k = 1_000_000, 1024, 1024, 10
c = DeviceCorpus(d, "f32", capacity=rows); c.fill_synthetic(seed=11, nrows=rows)
q_host = _lib.pinned_empty((B, d), np.float32); q_host[:] = synth.unit_queries(B, d, 12)
out = (_lib.pinned_empty((B, k), np.int32), _lib.pinned_empty((B, k), np.float64), _lib.pinned_empty((B,), np.int32))
def loop(n=40):
    for _ in range(3): c.topk(q_host, k, out=out)
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); c.topk(q_host, k, out=out); t.append(1e3 * (time.perf_counter() - t0))
    return {"mean": round(float(np.mean(t)), 4), "median": round(float(np.median(t)), 4), "max": round(float(max(t)), 4)}
print(json.dumps({"no_sampler": loop()}), flush=True)
with bench.ClockSampler(0) as s:
    r = loop()
print(json.dumps({"with_sampler": r, "samples": len(s.sm)}), flush=True)
