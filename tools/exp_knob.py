#!/usr/bin/env python
"""A/B of the sample fraction knob on one box (alternating, same session)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    sys.path.insert(0, p)
import torch
from b200rag import DeviceCorpus, _lib, synth

def run(rows, dtype, B, k, divs, reps=6):
    c = DeviceCorpus(1024, dtype, capacity=rows)
    c.fill_synthetic(seed=11, nrows=rows)
    q = synth.unit_queries(B, 1024, 12)
    qd = torch.from_numpy(q).cuda()
    o_r = torch.empty((B, k), dtype=torch.int32, device="cuda"); o_s = torch.empty((B, k), dtype=torch.float64, device="cuda"); o_c = torch.empty((B,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    res = {d: [] for d in divs}
    for rep in range(reps):
        for d in divs:
            _lib.set_option(KNOB, d)
            for _ in range(2):
                c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
            t = []
            for _ in range(3):             # the call is stream-ordered: time 4 of them with CUDA events on its stream
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st = STREAM
                e0.record(st)
                for _ in range(4):
                    c.topk_dev(qd.data_ptr(), B, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr())
                e1.record(st)
                torch.cuda.synchronize()
                t.append((e0.elapsed_time(e1) / 4, float(_lib.last_timings()[0])))
            res[d].append(min(t))
    _lib.set_option(KNOB, 1)
    out = {d: (round(float(np.median([x[0] for x in v])), 4), round(float(np.median([x[1] for x in v])), 4)) for d, v in res.items()}
    print(json.dumps({"rows": rows, "dtype": dtype, "B": B, "k": k, "call_ms,contraction_ms by knob value": out, "fallbacks": _lib.counters()["fallbacks"]}), flush=True)
    c.close()

STREAM = torch.cuda.Stream()
_lib.lib()
_lib.set_stream(STREAM.cuda_stream)
torch.cuda.set_stream(STREAM)
KNOB = sys.argv[1] if len(sys.argv) > 1 else "sample_div"
VALS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
run(1_000_000, "f32", 1024, 10, VALS)
run(1_250_000, "bf16", 4096, 100, VALS)
run(1_000_000, "f32", 256, 10, VALS)
