"""Worker of tests/test_gpu.py::test_two_rank_peer_exchange_vs_oracle: one process per GPU (torch.distributed.run,
NCCL for the plumbing), every rank holds a row shard in a DeviceCorpus; the merged result of the stream-ordered
device step (peer-memory exchange over NVLink: P2P stores + epoch flags + rank merge) and of the host-buffer call
is compared with the oracle over the WHOLE corpus on every rank.  Different (B, k) per case, one rank delayed on
purpose (skew).  Exit code 0 = all ranks equal to the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rag-dpo_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import helpers
    from b200rag import _lib
    from b200rag.sharded import ShardedDenseIndex
    from oracle import c_oracle
    from oracle import numpy_oracle as no

    _lib.lib()
    stream = torch.cuda.Stream(device=dev)
    _lib.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)
    n, d = 30_011, 256                                  # ragged shards
    failures = []
    for dtype, odt in (("f32", no.DT_F32), ("bf16", no.DT_BF16)):
        index = ShardedDenseIndex(d, n, dtype=dtype, device=dev)
        index.fill_synthetic(41)
        # the whole corpus on every rank, for the oracle: gather the shards (fp32 values that are exactly the stored ones)
        mine = index.corpus.download()
        per = (n + world - 1) // world
        buf = np.zeros((per, d), dtype=np.float32)
        buf[: len(mine)] = mine
        t_all = torch.empty((world, per, d), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(t_all, torch.from_numpy(buf).to(dev))
        full32 = t_all.cpu().numpy().reshape(world * per, d)[:n]
        full = (np.ascontiguousarray(full32, np.float32) if odt == no.DT_F32 else no.f32_to_bf16_bits(full32))
        for case, (B, k) in enumerate([(5, 10), (130, 100), (1, 1), (64, 50), (257, 10)]):
            q = helpers.synth_unit(B, d, seed=100 + case)
            er, es, ec = c_oracle.dense_topk(q, full, odt, k)
            qd = torch.from_numpy(q).to(dev)
            step, out = index.make_device_step(qd.data_ptr(), B, k)
            if index.exchange(B, k) is None:
                failures.append(f"{dtype} B={B} k={k}: the peer-memory exchange is not in use")
            for rep in range(3):                        # several epochs: both buffer parities
                if rank == (case + rep) % world:
                    time.sleep(0.02)                    # skew: this rank arrives late
                step()
                torch.cuda.synchronize()
                ids = out["ids"].cpu().numpy()
                scores = out["scores"].cpu().numpy()
                counts = out["counts"].cpu().numpy()
                if not (ids.tolist() == er.tolist() and np.array_equal(scores, es) and counts.tolist() == ec.tolist()):
                    failures.append(f"{dtype} device step B={B} k={k} rep={rep}: merged result differs from the oracle")
            h_ids, h_scores, h_counts = index.topk(q, k)
            if not (np.asarray(h_ids).tolist() == er.tolist() and np.array_equal(h_scores, es)):
                failures.append(f"{dtype} host call B={B} k={k}: merged result differs from the oracle")
        index.close()
        index.corpus.close()
    bad = torch.tensor([float(len(failures))], device=dev)
    dist.all_reduce(bad)
    if failures:
        print(f"rank {rank}: " + "; ".join(failures), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if bad.item() > 0 else 0)


if __name__ == "__main__":
    main()
