"""world_size-2 gloo test (CPU) of the row-sharded path: shard plan, global ids, all-gather layout and the
(score desc, id asc) merge.  The local top-k is supplied by the oracle here (the product path needs a GPU);
on the GPU box the same class runs with DeviceCorpus shards + rag_merge_topk_dev (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, d, k, out_dir):
    for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers
    from b200rag.sharded import ShardedDenseIndex, shard_bounds
    from oracle import c_oracle
    from oracle import numpy_oracle as no

    x = helpers.synth_unit(n, d, seed=5)
    x[n - 3] = x[2]                      # a tie across shards: lowest global id must win
    q = helpers.synth_unit(6, d, seed=6)
    q[0] = x[2]
    lo, hi = shard_bounds(n, world, rank)

    def local_topk(q32, kk):
        kl = min(kk, hi - lo)
        rows = np.full((len(q32), kk), -1, np.int32)
        scores = np.zeros((len(q32), kk), np.float64)
        counts = np.zeros(len(q32), np.int32)
        if kl > 0:
            r, s, c = c_oracle.dense_topk(q32, x[lo:hi], no.DT_F32, kl)
            rows[:, :kl], scores[:, :kl], counts[:] = r, s, c
        return rows, scores, counts

    index = ShardedDenseIndex(d, n, group=None, local_topk=local_topk)
    assert (index.row_lo, index.row_hi) == (lo, hi)
    ids, scores, counts = index.topk(q, k)
    er, es, ec = c_oracle.dense_topk(q, x, no.DT_F32, k)
    ok = ids.tolist() == er.tolist() and np.array_equal(scores, es) and counts.tolist() == ec.tolist()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else f"MISMATCH\n{ids.tolist()}\n{er.tolist()}")
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(501, 10), (64, 50), (7, 10)])
def test_sharded_topk_gloo_world2(tmp_path, n, k):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 64, k, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
