#!/usr/bin/env python
"""What one tcgen05.mma shape sustains on this GPU with nothing else going on: every SM multiplies resident
shared-memory operands (cbk_selftest_umma_rate).  Places the tensor-bound kernels (exhaustive MaxSim) against the
MMA shape's own ceiling rather than against cuBLAS's 256-wide 2-CTA tiles.

    python benchmarks/umma_rate.py [--iters 20000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20000)
    args = ap.parse_args()
    import torch
    from colbert_b200 import kernels
    dev = torch.device("cuda:0")
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    rows = []
    for (N, mode, n_acc, cps) in [(128, 0, 1, 1), (128, 0, 4, 1), (256, 0, 2, 1), (64, 0, 4, 1), (128, 1, 2, 1), (256, 1, 1, 1),
                                  (128, 0, 2, 2), (64, 0, 2, 4),
                                  (128, 4, 1, 1), (128, 4, 4, 1), (64, 4, 4, 1), (256, 4, 2, 1), (128, 4, 2, 2), (128, 20, 4, 1), (256, 20, 2, 1), (128, 6, 4, 1), (128, 22, 4, 1), (64, 6, 4, 1)]:
        kernels.selftest_umma_rate(N, mode, 200, n_acc, cps, dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cyc = kernels.selftest_umma_rate(N, mode, args.iters, n_acc, cps, dev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        flop = 2.0 * 128 * N * 128 * args.iters * n_sm * cps
        c = cyc.float()
        rows.append({"N": N, "A": "tmem" if mode & 1 else "smem", "issue": ("two warps, elect_one each" if mode & 2 else "elect_one in a converged warp") if mode & 4 else "thread 0 in a divergent branch",
                     "concurrent_tmem_readers": 16 if mode & 16 else 0, "accumulators": n_acc, "ctas_per_sm": cps,
                     "cycles_per_tile_per_cta": round(float(c.mean()) / args.iters, 1),
                     "cycles_per_tile_per_sm": round(float(c.mean()) / args.iters / cps, 1),
                     "ms": round(ms, 3), "pflops": round(flop / ms / 1e12, 3)})
        print(json.dumps(rows[-1]), flush=True)
    # TMEM read rate: warps per CTA x loads in flight
    for warps, depth in [(4, 1), (4, 2), (8, 2), (16, 1), (16, 2)]:
        it = 4000
        kernels.selftest_umma_rate(warps, 8, 64, depth, 1, dev)
        torch.cuda.synchronize()
        cyc = kernels.selftest_umma_rate(warps, 8, it, depth, 1, dev).float()
        torch.cuda.synchronize()
        per_ld = float(cyc.mean()) / it
        print(json.dumps({"tmem_read": True, "warps": warps, "loads_between_waits": depth,
                          "cycles_per_4KB_load_per_warp": round(per_ld, 1),
                          "bytes_per_cycle_per_sm": round(4096.0 * warps / per_ld, 1)}), flush=True)


if __name__ == "__main__":
    main()
