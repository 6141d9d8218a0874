#!/usr/bin/env python
"""Turns the JSON lines of benchmarks/exhaustive_sweep.py and benchmarks/umma_rate.py into profiles/rNN_exhaustive_sweep.md.

    python benchmarks/make_sweep_report.py gpurun_out/sweep_main.jsonl gpurun_out/sweep_fixed.jsonl gpurun_out/sweep_bf16.jsonl \\
        gpurun_out/umma_rate.jsonl > profiles/r01_exhaustive_sweep.md
"""
import json
import sys


def rows(fn):
    out = []
    for line in open(fn):
        line = line.strip()
        if line.startswith("{"):
            out.append(json.loads(line))
    return out


def line(r):
    return (f"| {r['doclen']} | {r['nq']} | {r['ms_score']} | {r['ms_topk']} | {r['hbm_gbs']} | {r['hbm_frac']} | {r['tflops']} | "
            f"{r['tensor_frac_of_sustained']} |\n")


def main():
    main_rows, fixed, bf16, rate = (rows(f) for f in sys.argv[1:5])
    out = """# Round 1 — query-batched exhaustive MaxSim (configs[3] shard, configs[4] sweep), 1×B200

Command: `python benchmarks/exhaustive_sweep.py --nq 1,2,4,8,16,32,64,128` (config 4: one GPU's shard of the 8.8 M-passage
corpus at 8 GPUs = 1.1 M docs, doclen U[20,120], 77.0 M rows, 19.7 GB) and
`python benchmarks/exhaustive_sweep.py --docs 300000 --doclen L --nq 1,2,4,8,16,64,256 --iters 3` for L in 32…512 (config 5).
CUDA-event timing of `cbk_maxsim_exhaustive` (ms_score) and `cbk_topk_dense` with k=1000 (ms_topk), after 2 warm-up calls.
`HBM GB/s` = store bytes ÷ ms_score (the store is read once per pass of 16 queries: above Nq=16 the kernel re-reads it, so this
column is only meaningful on the HBM-bound side); `TFLOP/s` = 2·32·Nq·128·rows ÷ ms_score.  Peaks: MEASURED_PEAKS.json
(6551 GB/s copy, 1387.7 TFLOP/s sustained bf16).

This is the kernel as committed at the end of round 1 (TMA and tcgen05.mma issued through `elect.sync` on converged warps,
two MMA-issuing warps with hoisted descriptors, out-of-line document-close path).  The table from the middle of the round
(same commands, one issuer in a `lane == 0` branch) read 3.46 ms at Nq=1 and 12.65 ms at Nq=16 for the U[20,120] shard.

## fp16 store

| doclen | Nq | ms score | ms top-k | HBM GB/s | of HBM peak | TFLOP/s | of sustained tensor peak |
|---|---:|---:|---:|---:|---:|---:|---:|
"""
    out += "".join(line(r) for r in main_rows + fixed)
    out += """
## bf16 store (query split in bf16 hi + lo parts, two MMA passes per tile, 8 queries per pass)

| doclen | Nq | ms score | ms top-k | HBM GB/s | of HBM peak | TFLOP/s (useful) | of sustained tensor peak |
|---|---:|---:|---:|---:|---:|---:|---:|
"""
    out += "".join(line(r) for r in bf16)
    out += """
## What one MMA shape sustains (`python benchmarks/umma_rate.py`, same box)

Every SM multiplies resident shared-memory operands, K = 128 per tile (8 × `tcgen05.mma` 128×N×16), nothing else running.

| N | A operand | issue | TMEM readers | accumulators | CTAs/SM | cycles per tile per SM | PFLOP/s |
|---:|---|---|---:|---:|---:|---:|---:|
"""
    for r in rate:
        if r.get("tmem_read"):
            continue
        out += (f"| {r['N']} | {r['A']} | {r['issue']} | {r['concurrent_tmem_readers']} | {r['accumulators']} | {r['ctas_per_sm']} | "
                f"{r['cycles_per_tile_per_sm']} | {r['pflops']} |\n")
    out += """
Reading: (1) a `tcgen05.mma` issued by `threadIdx.x == 0` inside a divergent branch costs ≈ 154 cycles whatever N is, because
ptxas wraps it in an ELECT / R2UR / branch loop; issued by `elect.sync` on a converged warp it is a bare `UTCHMMA`.  (2) One
issuing warp then sustains one 128×128×16 MMA per 86–90 cycles (floor 64) and one 128×256×16 per 128 cycles (= floor);
TWO issuing warps — in two CTAs or in one — reach the 64-cycle floor with N = 128 (2.2 PFLOP/s).  (3) Sixteen warps reading
the accumulators back with `tcgen05.ld` cost the MMA stream 5–12 %.  The exhaustive kernel (N = 128, two issuers, four groups
draining while the next accumulators are multiplied) reaches 0.96–1.15 PFLOP/s; what is left is the single-buffered
accumulator hand-off (TMEM holds 4 × 128 columns = one accumulator per query block of a 16-query pass): ncu shows the
epilogue warps waiting for their next accumulator 41 % of the time.

TMEM read rate (`tcgen05.ld.32x32b.x32` + a 32-wide max tree per thread):

| warps | loads between waits | cycles per 4 KB load per warp | B/cycle/SM |
|---:|---:|---:|---:|
"""
    for r in rate:
        if r.get("tmem_read"):
            out += f"| {r['warps']} | {r['loads_between_waits']} | {r['cycles_per_4KB_load_per_warp']} | {r['bytes_per_cycle_per_sm']} |\n"
    sys.stdout.write(out)


if __name__ == "__main__":
    main()
