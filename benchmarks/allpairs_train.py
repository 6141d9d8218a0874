#!/usr/bin/env python
"""Training-shaped all-pairs MaxSim: BaseModel.score forward + backward at the author's batch
(colbert_model.py:87-95 on the all-gathered batch: q = 170 questions x m = 32, d = 340 passages x n = 384, h = 768).

    python benchmarks/allpairs_train.py [--q 170 --m 32 --d 340 --n 384 --h 768] [--iters 20] [--torch-ref]

Prints one JSON line: forward kernel ms and TFLOP/s (2·q·m·d·n·h) against the sustained cuBLAS bf16 figure of
MEASURED_PEAKS.json, the mask+cast pass, the backward pass, and — with --torch-ref — the reference's own op sequence
(einsum → max → sum, fp32 and bf16-autocast) in eager PyTorch on the same GPU, chunked over queries so that `simmat` fits."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=170)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--d", type=int, default=340)
    ap.add_argument("--n", type=int, default=384)
    ap.add_argument("--h", type=int, default=768)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--torch-ref", action="store_true")
    args = ap.parse_args()
    import torch
    from colbert_b200 import kernels
    from colbert_b200.modeling.BaseModel import BaseModel
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    Q = torch.nn.functional.normalize(torch.randn(args.q, args.m, args.h, device=dev), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(args.d, args.n, args.h, device=dev), dim=-1)
    qmask = torch.ones(args.q, args.m, dtype=torch.int64, device=dev)
    dmask = (torch.arange(args.n, device=dev)[None, :] < torch.randint(args.n // 4, args.n + 1, (args.d, 1), device=dev)).long()
    W = torch.randn(args.q, args.d, device=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    flops = 2.0 * args.q * args.m * args.d * args.n * args.h

    Qp = kernels.mask_cast_rows(Q.reshape(-1, args.h), qmask.reshape(-1), torch.float16).reshape(args.q, args.m, args.h)
    Dp = kernels.mask_cast_rows(D.reshape(-1, args.h), dmask.reshape(-1), torch.float16).reshape(args.d, args.n, args.h)
    scores, argmax = kernels.score_allpairs_fwd(Qp, Dp)
    ms_cast = timed(lambda: (kernels.mask_cast_rows(Q.reshape(-1, args.h), qmask.reshape(-1), torch.float16),
                             kernels.mask_cast_rows(D.reshape(-1, args.h), dmask.reshape(-1), torch.float16)), args.iters)
    ms_fwd = timed(lambda: kernels.score_allpairs_fwd(Qp, Dp), args.iters)
    ms_fwd_noarg = timed(lambda: kernels.score_allpairs_fwd(Qp, Dp, want_argmax=False), args.iters)
    qm, dm = qmask.reshape(-1), dmask.reshape(-1)
    ms_bwd = timed(lambda: kernels.score_allpairs_bwd(Qp, Dp, W, argmax, qm, dm), args.iters)
    ms_dq = timed(lambda: kernels.score_allpairs_bwd(Qp, Dp, W, argmax, qm, dm, True, False), args.iters)
    ms_dd = timed(lambda: kernels.score_allpairs_bwd(Qp, Dp, W, argmax, qm, dm, False, True), args.iters)

    def step():
        Qa, Da = Q.detach().requires_grad_(True), D.detach().requires_grad_(True)
        (BaseModel.score(Qa, Da, qmask, dmask) * W).sum().backward()
    ms_step = timed(step, args.iters)
    out = {"shape": {"q": args.q, "m": args.m, "d": args.d, "n": args.n, "h": args.h}, "gflop_forward": flops / 1e9,
           "ms_mask_cast": round(ms_cast, 4), "ms_forward_kernel": round(ms_fwd, 4), "ms_forward_kernel_no_argmax": round(ms_fwd_noarg, 4),
           "forward_tflops": round(flops / (ms_fwd * 1e-3) / 1e12, 1), "forward_frac_of_sustained_peak": round(flops / (ms_fwd * 1e-3) / 1e12 / tf_peak, 3),
           "ms_backward": round(ms_bwd, 4), "ms_backward_dQ": round(ms_dq, 4), "ms_backward_dD": round(ms_dd, 4),
           "ms_score_fwd_bwd_through_autograd": round(ms_step, 4)}

    if args.torch_ref:
        def ref_step(dtype):
            Qa, Da = Q.detach().requires_grad_(True), D.detach().requires_grad_(True)
            total = 0.0
            for lo in range(0, args.q, 16):   # simmat [16, d, m, n] fp32 = 267 MB per chunk at the default shape
                with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
                    sim = torch.einsum("qmh,dnh->qdmn", Qa[lo:lo + 16] * qmask[lo:lo + 16, :, None], Da * dmask[..., None])
                s = sim.float().max(-1)[0].sum(-1)
                total = total + (s * W[lo:lo + 16]).sum()
            total.backward()
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        out["ms_torch_eager_fp32_fwd_bwd"] = round(timed(lambda: ref_step(None), 3, 1), 3)
        torch.backends.cuda.matmul.allow_tf32 = True
        out["ms_torch_eager_tf32_fwd_bwd"] = round(timed(lambda: ref_step(None), 3, 1), 3)
        torch.backends.cuda.matmul.allow_tf32 = prev
        out["ms_torch_eager_bf16_autocast_fwd_bwd"] = round(timed(lambda: ref_step(torch.bfloat16), 3, 1), 3)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
