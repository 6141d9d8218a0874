#!/usr/bin/env python
"""BASELINE.json configs[3]: exhaustive MaxSim over a corpus sharded by document range across the GPUs of one box, top-1000
merged through one NCCL all-gather of packed keys.  Weak scaling in the corpus: every GPU holds `--docs-per-gpu` passages
(1.1 M × ~70 tokens × 128-d fp16 = 19.7 GB, i.e. the 8.8 M-passage corpus at 8 GPUs); every query is scored against ALL of
them.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \\
        benchmarks/exhaustive_sharded.py [--nq 16] [--k 1000] [--docs-per-gpu 1100000] [--iters 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs-per-gpu", type=int, default=1_100_000)
    ap.add_argument("--nq", default="1,16,64")
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from colbert_b200.ranking import ColbertRanker
    from colbert_b200.sharding import ShardedColbertRanker
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(1000 + rank)
    doclens = torch.from_numpy(rng.integers(20, 121, size=args.docs_per_gpu).astype(np.int64))
    n_tok = int(doclens.sum())
    store = torch.empty((n_tok + 512, 128), dtype=torch.float16, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(77 + rank)
    for lo in range(0, n_tok + 512, 1 << 22):
        hi = min(lo + (1 << 22), n_tok + 512)
        x = torch.randn((hi - lo, 128), generator=g, device=dev)
        store[lo:hi] = torch.nn.functional.normalize(x, dim=1).to(torch.float16)
    store[n_tok:] = 0
    local = ColbertRanker.from_store(store, doclens)
    strides = [45, 70, 95, 120]                       # the whole corpus' percentile strides (same law on every shard)
    sharded = ShardedColbertRanker(local, rank * args.docs_per_gpu, strides) if world > 1 else None
    if sharded is None:
        local.strides = strides
    gq = torch.Generator(device="cpu")
    gq.manual_seed(5)
    for nq in [int(x) for x in args.nq.split(",")]:
        Q = torch.nn.functional.normalize(torch.randn((nq, 32, 128), generator=gq), dim=2).to(dev)   # replicated queries
        run = (lambda: sharded.rank_exhaustive(Q, args.k)) if sharded is not None else (lambda: local.rank_exhaustive(Q, args.k))
        for _ in range(2):
            pids, scores = run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            pids, scores = run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], dtype=torch.float64, device=dev)
        tok = torch.tensor([float(n_tok)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tok, op=dist.ReduceOp.SUM)
            ref = pids.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, pids), "ranks disagree on the merged top-k"
        ms, total_tok = float(t), float(tok)
        if rank == 0:
            assert bool((scores[:, :-1] >= scores[:, 1:]).all())
            print(json.dumps({"config": "exhaustive, store sharded by document range", "n_gpus": world, "nq": nq, "k": args.k,
                              "docs_total": args.docs_per_gpu * world, "tokens_total": int(total_tok),
                              "store_gb_total": round(total_tok * 256 / 1e9, 1), "ms_per_batch": round(ms, 3),
                              "queries_per_s": round(nq / ms * 1e3, 1), "docs_scored_per_s": round(nq * args.docs_per_gpu * world / ms * 1e3),
                              "tflops_total": round(2 * 32 * nq * 128 * total_tok / ms / 1e9, 1),
                              "hbm_gbs_per_gpu": round(total_tok / world * 256 / ms / 1e6 * -(-nq // 16) , 1)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
