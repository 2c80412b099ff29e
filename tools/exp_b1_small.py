#!/usr/bin/env python
"""single-query dense latency on small fp32 corpora: bf16-shadow contraction path vs the fp32 scan path (option
tc_b1_shadow), device time per call with CUDA events over back-to-back stream-ordered calls"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200")):
    sys.path.insert(0, p)
import torch
from b200rag import DeviceCorpus, _lib, synth

_lib.lib()
STREAM = torch.cuda.Stream()
_lib.set_stream(STREAM.cuda_stream)
torch.cuda.set_stream(STREAM)
for rows in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16919,50000,200000,1000000").split(",")]:
    c = DeviceCorpus(1024, "f32", capacity=rows)
    c.fill_synthetic(seed=11, nrows=rows)
    q = torch.from_numpy(synth.unit_queries(1, 1024, 12)).cuda()
    k = 50
    o = (torch.empty((1, k), dtype=torch.int32, device="cuda"), torch.empty((1, k), dtype=torch.float64, device="cuda"),
         torch.empty((1,), dtype=torch.int32, device="cuda"))
    res = {}
    for shadow in (1, 0, 1, 0):
        _lib.set_option("tc_b1_shadow", shadow)
        for _ in range(5):
            c.topk_dev(q.data_ptr(), 1, k, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(STREAM)
        for _ in range(50):
            c.topk_dev(q.data_ptr(), 1, k, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr())
        e1.record(STREAM)
        torch.cuda.synchronize()
        res.setdefault(shadow, []).append(round(e0.elapsed_time(e1) / 50, 4))
    _lib.set_option("tc_b1_shadow", 1)
    print(json.dumps({"rows": rows, "k": k, "device_ms_per_call shadow(1)/scan(0)": res}), flush=True)
    c.close()
