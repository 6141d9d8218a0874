#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[3] and [4]): query-batched exhaustive MaxSim over one GPU's
shard of the 8.8 M-passage corpus, and the Nq x doclen sweep showing the HBM-bound → tensor-bound transition.

    python benchmarks/exhaustive_sweep.py [--docs 1100000] [--nq 1,4,8,16,32,64,128] [--doclen 0|32|64|...] [--dtype fp16]

Prints one JSON line per (Nq, doclen) point: ms per call, docs scored/s, algorithmic GB/s (store bytes read
once per pass of 16 queries are NOT what is counted: algorithmic = store bytes once per call) and TFLOP/s
(2*32*Nq*128*tokens), each with its fraction of the measured peak (MEASURED_PEAKS.json)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_100_000)       # 8.8 M / 8 GPUs
    ap.add_argument("--nq", default="1,4,8,16,32,64,128")
    ap.add_argument("--doclen", type=int, default=0, help="0: U[20,120] (mean 70, config 4); else fixed (config 5)")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--dim", type=int, default=128, help="embedding width; other than 128 needs a fixed --doclen and fp16 "
                                                         "(multi-view store scored by the all-pairs kernel)")
    ap.add_argument("--q-len", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import torch
    from colbert_b200 import kernels
    from colbert_b200.ranking import ColbertRanker
    dev = torch.device("cuda", 0)
    dt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    g = torch.Generator().manual_seed(7)
    if args.doclen:
        doclens = torch.full((args.docs,), args.doclen, dtype=torch.int64)
    else:
        doclens = torch.randint(20, 121, (args.docs,), generator=g, dtype=torch.int64)
    total = int(doclens.sum())
    store = torch.zeros(total + 512, args.dim, dtype=dt, device=dev)
    gg = torch.Generator(device=dev).manual_seed(8)
    for s in range(0, total, 1 << 22):
        e = min(total, s + (1 << 22))
        store[s:e] = torch.nn.functional.normalize(torch.randn(e - s, args.dim, generator=gg, device=dev), dim=1).to(dt)
    ranker = ColbertRanker.from_store(store, doclens)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak, tf_peak = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    for nq in [int(x) for x in args.nq.split(",")]:
        Q = torch.nn.functional.normalize(torch.randn(nq, args.q_len, args.dim, generator=g), dim=2).to(dev)
        out = torch.empty((nq, args.docs), dtype=torch.float32, device=dev)
        for _ in range(2):
            ranker.score_all(Q)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(args.iters):
            sc = ranker.score_all(Q)
        e1.record()
        for _ in range(args.iters):
            kernels.topk_dense(sc, min(args.k, args.docs))
        e2.record()
        torch.cuda.synchronize()
        ms, ms_topk = e0.elapsed_time(e1) / args.iters, e1.elapsed_time(e2) / args.iters
        gbs = total * args.dim * 2 / (ms * 1e-3) / 1e9
        tfl = 2.0 * args.q_len * nq * args.dim * total / (ms * 1e-3) / 1e12
        print(json.dumps({"nq": nq, "doclen": args.doclen or "U[20,120]", "dtype": args.dtype, "docs": args.docs,
                          "dim": args.dim, "q_len": args.q_len, "tokens": total, "store_gb": total * args.dim * 2 / 1e9, "ms_score": round(ms, 3), "ms_topk": round(ms_topk, 3),
                          "docs_scored_per_s": nq * args.docs / ((ms + ms_topk) * 1e-3),
                          "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / hbm_peak, 3),
                          "tflops": round(tfl, 1), "tensor_frac_of_sustained": round(tfl / tf_peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
