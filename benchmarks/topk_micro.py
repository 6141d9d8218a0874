#!/usr/bin/env python
"""Per-query top-10 over short candidate lists at the shapes the rerank step produces: 4096 x 1000 (one GPU), 8192 x 500 and
32768 x 125 (the routed lists of 2- and 8-GPU shards).  CUDA-event time of cbk_topk_per_query_keys.

    python benchmarks/topk_micro.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from colbert_b200 import _lib, kernels
    dev = torch.device("cuda:0")
    for B, n in ((4096, 1000), (8192, 500), (32768, 125)):
        sc = torch.randn(B * n, device=dev)
        ids = torch.randint(0, 1 << 30, (B * n,), device=dev, dtype=torch.int64)
        rp = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=dev)
        for _ in range(3):
            kernels.topk_per_query(sc, ids, rp, 10, 1000, flags=_lib.CBK_TOPK_NEG_INF_IS_PADDING, as_keys=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            kernels.topk_per_query(sc, ids, rp, 10, 1000, flags=_lib.CBK_TOPK_NEG_INF_IS_PADDING, as_keys=True)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"queries": B, "candidates_per_query": n, "k": 10, "ms": round(e0.elapsed_time(e1) / 20, 4)}), flush=True)


if __name__ == "__main__":
    main()
