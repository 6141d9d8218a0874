#!/usr/bin/env python
"""Latency of the reference-shaped call: ONE query, 1000 candidates, `ColbertRanker.rank_forward(Q, pids, depth=10)`
(reference colbert_ranker.py:75-137), Python lists in and out (configs[0] shape).  The CPU side of the comparison is
`python bench.py --impl reference` (the port of the reference's ranker on the host cores: 0.31 M candidates/s = 3.2 ms per
1000-candidate call on the 16-core box).

    python benchmarks/single_query_latency.py [--docs 200000] [--iters 200]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=200_000)
    ap.add_argument("--cands", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    import numpy as np
    import torch
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    index = synthetic.make_index(3, args.docs, dim=128, lo=1, hi=180)
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device="cuda:0")
    Q = synthetic.make_queries(4, 64, 32, 128)
    cand = synthetic.make_candidates(5, 64, index.num_docs, args.cands)
    Qs = [torch.from_numpy(Q[i]).unsqueeze(0).permute(0, 2, 1) for i in range(64)]
    cl = [c.tolist() for c in cand]
    for i in range(10):
        ranker.rank_forward(Qs[i], cl[i], depth=10)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.iters):
        ranker.rank_forward(Qs[i % 64], cl[i % 64], depth=10)          # returns Python lists ⇒ synchronises
    gpu_ms = (time.perf_counter() - t0) / args.iters * 1e3
    # same call, query in the reference's own dim-major layout (transposed while staging)
    Qc = [q.contiguous() for q in Qs]
    t0 = time.perf_counter()
    for i in range(args.iters):
        ranker.rank_forward(Qc[i % 64], cl[i % 64], depth=10)
    gpu_dm_ms = (time.perf_counter() - t0) / args.iters * 1e3
    # the general path (device tensors, separate torch copies + two library calls), for comparison
    Qg = [q.to(ranker.device) for q in Qs]
    for i in range(5):
        ranker.rank_forward(Qg[i], cl[i], depth=10)
    t0 = time.perf_counter()
    for i in range(args.iters):
        ranker.rank_forward(Qg[i % 64], cl[i % 64], depth=10)
    gpu_general_ms = (time.perf_counter() - t0) / args.iters * 1e3
    # kernel-only time of the same work (device tensors prepared once)
    dev = ranker.device
    Qd = torch.from_numpy(Q[0:1]).to(dev)
    cd = torch.from_numpy(cand[0]).to(dev)
    rp = torch.tensor([0, args.cands], dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        ranker.score_candidates(Qd, cd, rp)
    e0.record()
    for _ in range(args.iters):
        ranker.score_candidates(Qd, cd, rp)
    e1.record()
    torch.cuda.synchronize()
    kern_ms = e0.elapsed_time(e1) / args.iters
    # the CUDA-graph low-batch entry (H2D + MaxSim + top-k + D2H replayed as one graph): 1 and 8 queries per call
    from colbert_b200.ranking.pipeline import GraphedRerank
    graph_ms = {}
    for b in (1, 8):
        g = GraphedRerank(ranker, batch=b, q_len=32, n_cand=args.cands, depth=10)
        Qb = torch.from_numpy(Q[:b])
        Cb = torch.from_numpy(np.stack([cand[i] for i in range(b)]))
        for _ in range(10):
            g(Qb, Cb)
        t0 = time.perf_counter()
        for _ in range(args.iters):
            g(Qb, Cb)
        graph_ms[b] = (time.perf_counter() - t0) / args.iters * 1e3
    print(json.dumps({"call": "rank_forward(Q[1,128,32], 1000 pids, depth=10)", "gpu_ms_per_call": round(gpu_ms, 4),
                      "graphed_rerank_ms_per_call_batch1": round(graph_ms[1], 4),
                      "graphed_rerank_ms_per_call_batch8": round(graph_ms[8], 4),
                      "gpu_ms_per_call_dim_major_q": round(gpu_dm_ms, 4), "gpu_ms_per_call_general_path": round(gpu_general_ms, 4),
                      "gpu_maxsim_kernel_ms": round(kern_ms, 4),
                      "cpu_reference": "python bench.py --impl reference  (candidates/s of the CPU port; 1000 / value = s per call)"}))


if __name__ == "__main__":
    main()
