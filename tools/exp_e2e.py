#!/usr/bin/env python
"""Where does the host-buffer call (rag_dense_topk) spend its time?  Config 2 batch 1024."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    sys.path.insert(0, p)
import torch
from b200rag import DeviceCorpus, _lib, synth

rows, d, B, k = 1_000_000, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 10
c = DeviceCorpus(d, "f32", capacity=rows)
c.fill_synthetic(seed=11, nrows=rows)
q_host = _lib.pinned_empty((B, d), np.float32)
q_host[:] = synth.unit_queries(B, d, 12)
out = (_lib.pinned_empty((B, k), np.int32), _lib.pinned_empty((B, k), np.float64), _lib.pinned_empty((B,), np.int32))
qd = torch.from_numpy(np.array(q_host)).cuda()
o_r = torch.empty((B, k), dtype=torch.int32, device="cuda"); o_s = torch.empty((B, k), dtype=torch.float64, device="cuda"); o_c = torch.empty((B,), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()

def med(f, n=30):
    for _ in range(3): f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); t.append(1e3 * (time.perf_counter() - t0))
    return round(float(np.median(t)), 4)

res = {}
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
res["mode"] = mode
sampler = None
if "stream" in mode:
    st = torch.cuda.Stream()
    _lib.set_stream(st.cuda_stream)
    torch.cuda.set_stream(st)
if "nvml" in mode:
    sys.path.insert(0, ROOT)
    import bench
    sampler = bench.ClockSampler(0)
    sampler.__enter__()
res["device_call_ms"] = med(lambda: c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr()))
res["device_stage_ms"] = [round(float(v), 4) for v in _lib.last_timings()[:4]]
res["host_call_ms"] = med(lambda: c.topk(q_host, k, out=out))
tm = _lib.last_timings()
res["host_stage_ms"] = [round(float(v), 4) for v in tm[:4]]
res["host_queue_ms,total_ms"] = [round(float(tm[4]), 4), round(float(tm[5]), 4)]
pin_t = torch.empty((B, d), dtype=torch.float32).pin_memory()
def h2d():
    qd.copy_(pin_t, non_blocking=True); torch.cuda.synchronize()
res["h2d_%dKB_sync_ms" % (B * d * 4 // 1024)] = med(h2d)
small = torch.empty((B, k), dtype=torch.float64).pin_memory()
def d2h():
    small.copy_(o_s, non_blocking=True); torch.cuda.synchronize()
res["d2h_small_sync_ms"] = med(d2h)
res["empty_sync_ms"] = med(lambda: torch.cuda.synchronize())
if sampler:
    sampler.__exit__()
    res["clock_samples"] = len(sampler.sm)
print(json.dumps(res), flush=True)
