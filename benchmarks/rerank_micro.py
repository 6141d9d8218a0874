#!/usr/bin/env python
"""Kernel-only timing of the rerank kernel on one workload shape (CUDA events, inputs resident in HBM):

    python benchmarks/rerank_micro.py --doclen 8 --q-len 8 [--dtype bf16] [--docs 2000000] [--queries 4096] [--cands 1000]
    CBK_RERANK_PROBE=1 python benchmarks/rerank_micro.py ...      # the gather alone: what the streaming structure sustains

--doclen 0 = U[1,180] (configs[1]); 8 / 16 = multi-view (configs[2]).  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--doclen", type=int, default=8)
    ap.add_argument("--q-len", type=int, default=8)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--docs", type=int, default=2_000_000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--cands", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--dim", type=int, default=128, help="embedding width (768: the author's un-projected multi-view index)")
    ap.add_argument("--ksplit", action="store_true", help="dim 192 … 1024: the K-split mma.sync kernel instead of the streaming one")
    ap.add_argument("--no-fixed", action="store_true", help="score through the looked-up path even on a fixed-doclen index")
    args = ap.parse_args()
    import torch
    import bench
    from colbert_b200 import _lib, kernels
    from colbert_b200.ranking import ColbertRanker
    dev = torch.device("cuda", 0)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    store, doclens = bench.build_store(torch, dev, args.docs, args.dim, dt, seed=777, doclen_fixed=args.doclen)
    ranker = ColbertRanker.from_store(store, doclens)
    g = torch.Generator().manual_seed(1)
    Q = torch.nn.functional.normalize(torch.randn(args.queries, args.q_len, args.dim, generator=g), dim=2).to(dev)
    cand = torch.randint(0, args.docs, (args.queries * args.cands,), generator=g, dtype=torch.int64).to(dev)
    rowptr = torch.arange(0, args.queries * args.cands + 1, args.cands, dtype=torch.int64, device=dev)
    flags = ranker.kernel_flags if args.no_fixed else ranker.effective_flags
    if args.ksplit:
        flags |= _lib.CBK_FLAG_RERANK_KSPLIT

    def run():
        return kernels.maxsim_rerank(ranker.tensor, ranker._pfxsum_dev, ranker._doclens_dev, ranker.strides, Q, cand, rowptr,
                                     flags=flags)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    algo = int(ranker._doclens_dev[cand].to(torch.int64).sum().item()) * args.dim * 2
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    print(json.dumps({"doclen": args.doclen or "U[1,180]", "dim": args.dim, "q_len": args.q_len, "dtype": args.dtype,
                      "fixed_path": bool(flags & _lib.CBK_FLAG_FIXED_DOCLEN), "ksplit": bool(args.ksplit), "probe_gather_only": bool(os.environ.get("CBK_RERANK_PROBE")),
                      "kernel_ms": round(ms, 4), "gbs": round(algo / ms / 1e6, 1), "frac_of_hbm_peak": round(algo / ms / 1e6 / peak, 3),
                      "cands_per_s": round(args.queries * args.cands / ms * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
