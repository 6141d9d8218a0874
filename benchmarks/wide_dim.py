#!/usr/bin/env python
"""Rerank throughput for embeddings wider than 128 (the author's configuration: dim = 768, reference proj_conf/dense.yaml:8),
as achieved HBM GB/s: the default kernel for the width (dim 192 … 1024: the tcgen05 streaming kernel, csrc/rerank_wide_stream.cu;
dim 64: the K-split mma.sync kernel, csrc/rerank_wide.cu), the K-split kernel forced (CBK_FLAG_RERANK_KSPLIT)
and the generic CUDA-core kernel.

    python benchmarks/wide_dim.py [--dims 768,1024,256] [--store-gb 10] [--queries 256] [--cands 1000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="768,1024,512,256,64")
    ap.add_argument("--store-gb", type=float, default=10.0)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--cands", type=int, default=1000)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    from colbert_b200 import _lib
    from colbert_b200.ranking import ColbertRanker
    dev = torch.device("cuda:0")
    dt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    peak = 6551.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    for dim in [int(x) for x in args.dims.split(",")]:
        rng = np.random.default_rng(dim)
        n_docs = int(args.store_gb * 1e9 / (90.5 * dim * 2))
        doclens = torch.from_numpy(rng.integers(1, 181, size=n_docs).astype(np.int64))
        n_tok = int(doclens.sum())
        store = torch.empty((n_tok + 512, dim), dtype=dt, device=dev)
        for lo in range(0, n_tok + 512, 1 << 20):                     # unit-norm random rows, chunked
            hi = min(lo + (1 << 20), n_tok + 512)
            x = torch.randn((hi - lo, dim), device=dev)
            store[lo:hi] = torch.nn.functional.normalize(x, dim=1).to(dt)
        store[n_tok:] = 0
        ranker = ColbertRanker.from_store(store, doclens)
        Q = torch.nn.functional.normalize(torch.randn((args.queries, 32, dim), device=dev), dim=2)
        cand = torch.from_numpy(rng.integers(0, n_docs, size=(args.queries, args.cands)).astype(np.int64)).to(dev)
        rowptr = torch.arange(0, (args.queries + 1) * args.cands, args.cands, dtype=torch.int64, device=dev)
        flat = cand.reshape(-1).contiguous()
        algo = float(doclens.to(dev)[flat].sum()) * dim * 2
        out = {"dim": dim, "dtype": args.dtype, "docs": n_docs, "store_gb": round(store.numel() * 2 / 1e9, 2),
               "candidates": flat.numel(), "algorithmic_gb": round(algo / 1e9, 2)}
        for name, flags, nq in (("default_kernel", 0, args.queries), ("k_split", _lib.CBK_FLAG_RERANK_KSPLIT, args.queries),
                                ("generic_cuda_core", _lib.CBK_FLAG_RERANK_GENERIC, max(1, args.queries // 16))):
            ranker.kernel_flags = flags
            f, rp, Qs = flat[: nq * args.cands], rowptr[: nq + 1], Q[:nq].contiguous()
            bytes_ = float(doclens.to(dev)[f].sum()) * dim * 2
            for _ in range(2):
                ranker.score_candidates(Qs, f, rp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                ranker.score_candidates(Qs, f, rp)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            out[name] = {"ms": round(ms, 3), "candidates_per_s": round(nq * args.cands / ms * 1e3), "hbm_gbs": round(bytes_ / ms / 1e6, 1),
                         "of_hbm_peak": round(bytes_ / ms / 1e6 / peak, 3)}
        out["default_vs_generic"] = round(out["default_kernel"]["hbm_gbs"] / out["generic_cuda_core"]["hbm_gbs"], 1)
        print(json.dumps(out), flush=True)
        del ranker, store
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
