#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native late-interaction scoring path.

Metric (BASELINE.json): candidates MaxSim-scored per second on the k=1000-candidate rerank workload
(configs[1]: 4,096 queries × 1,000 candidates, doclen ≤ 180, 128-d, bf16 store), whole job over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (pid→offset lookup + gather + MaxSim + per-query top-k) over one
batch of 4,096 queries.  `value` is measured with inputs resident in HBM; `e2e` is the same metric
through the public API (ColbertRanker.rank_forward_batch) with pinned HOST inputs and a host read of
the result inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidates MaxSim-scored/sec (k=1000 rerank)"
UNIT = "candidates/s"

# dram__bytes_read.sum + dram__bytes_write.sum of ONE maxsim_rerank_kernel launch of the default workload, from the
# `ncu --set full` capture summarised in profiles/r01_end_rerank_ncu_full_summary.csv (94.950 GB read + 0.019 GB
# written; algorithmic 94.869 GB — the difference is candidate/query metadata and DRAM sector granularity).  Only
# valid for the default arguments (same seeds → same candidate lists).
NCU_TRAFFIC_DEFAULT_BYTES = 94_968_766_096


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--cands", type=int, default=1000)
    ap.add_argument("--docs", type=int, default=2_000_000, help="documents per GPU shard")
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--doclen-fixed", type=int, default=0,
                    help="every document has exactly this many rows (configs[2], multi-view: 8); 0 = U[1,180]")
    ap.add_argument("--q-len", type=int, default=32, help="query rows (multi-view: q_view)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "mma", "tcgen05"],
                    help="rerank kernel: mma.sync (rerank.cu) or tcgen05/TMEM (rerank_umma.cu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dim", type=int, default=128, help="embedding width (configs use 128; the author's config is 768: "
                    "pair it with --docs 300000 so that the store fits)")
    ap.add_argument("--cpu-queries", type=int, default=128, help="bounded CPU-baseline sample (queries)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# synthetic workload: unit-norm rows, doclen U[1,180], candidates uniform w/o replacement
# --------------------------------------------------------------------------------------------------
def build_store(torch, dev, n_docs, dim, dtype, seed, doclen_fixed=0):
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    if doclen_fixed:
        doclens = torch.full((n_docs,), int(doclen_fixed), dtype=torch.int64)
    else:
        doclens = torch.randint(1, 181, (n_docs,), generator=g, dtype=torch.int64)
    total = int(doclens.sum())
    store = torch.zeros(total + 512, dim, dtype=dtype, device=dev)       # reference layout: +512 zero rows
    gg = torch.Generator(device=dev); gg.manual_seed(seed + 1)
    chunk = 1 << 22
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        x = torch.randn(e - s, dim, generator=gg, device=dev, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, p=2, dim=1)
        store[s:e] = x.to(dtype)
    return store, doclens


def build_queries(torch, n_q, q_len, dim, n_docs, n_cand, seed):
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    Q = torch.nn.functional.normalize(torch.randn(n_q, q_len, dim, generator=g), p=2, dim=2)
    # distinct pids per query, uniform over the corpus: draw a few extra, de-duplicate, shuffle back (the reference
    # receives list(set(...)), i.e. an arbitrary order)
    cand = torch.empty(n_q, n_cand, dtype=torch.int64)
    for b in range(0, n_q, 256):
        e = min(n_q, b + 256)
        draw = torch.randint(0, n_docs, (e - b, n_cand + 64), generator=g, dtype=torch.int64)
        for r in range(e - b):
            u = torch.unique(draw[r])                    # sorted unique
            u = u[torch.randperm(u.numel(), generator=g)][:n_cand]   # unsorted, as list(set(...)) order is arbitrary
            assert u.numel() == n_cand
            cand[b + r] = u
    return Q, cand


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's ranking algorithm (oracle/ref_port_torch.py) on the host cores
# --------------------------------------------------------------------------------------------------
class CpuArm:
    """CpuRankerPort.rank_forward on a bounded sample: a 50k-doc store (configs[0] shape), queries of
    32x128 fp32 with `n_cand` candidates each, all host threads."""

    def __init__(self, n_queries, n_cand, depth, seed=99):
        import torch
        from colbert_b200 import synthetic
        from oracle.ref_port_torch import CpuRankerPort
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        index = synthetic.make_index(seed, 50_000, dim=128, lo=1, hi=180)
        store = torch.zeros(index.num_tokens + 512, 128, dtype=torch.float16)
        store[: index.num_tokens] = torch.from_numpy(index.emb)
        self.port = CpuRankerPort(store, index.doclens.tolist(), max_candidates=n_cand)
        self.Q = torch.from_numpy(synthetic.make_queries(seed + 1, n_queries, 32, 128))
        cand = synthetic.make_candidates(seed + 2, n_queries, index.num_docs, n_cand)
        self.cand_lists = [c.tolist() for c in cand]
        self.n_queries, self.n_cand, self.depth, self.tokens = n_queries, n_cand, depth, index.num_tokens
        self.port.rank_forward(self.Q[0].unsqueeze(0).permute(0, 2, 1), self.cand_lists[0], depth=depth)  # warm-up

    def run(self):
        """one pass over the sample → seconds"""
        t0 = time.perf_counter()
        for b in range(self.n_queries):
            self.port.rank_forward(self.Q[b].unsqueeze(0).permute(0, 2, 1), self.cand_lists[b], depth=self.depth)
        return time.perf_counter() - t0

    def describe(self, extra=""):
        return (f"{self.n_queries} queries x {self.n_cand} candidates, Q 32x128 fp32, doclen U[1,180], fp16 store of "
                f"50k docs ({self.tokens} tokens), top-{self.depth}; oracle/ref_port_torch.py (the reference's op "
                f"sequence on torch CPU ops), {self.cores} threads{extra}")


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    per_step_q = 32
    arm = CpuArm(per_step_q, args.cands, args.depth)
    for _ in range(warmup):
        arm.run()
    total = sum(arm.run() for _ in range(steps))
    value = steps * per_step_q * args.cands / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"rerank {per_step_q}x{args.cands} per step (bounded sample of configs[1]: "
                               f"{args.queries}x{args.cands}, doclen<=180, dim 128)", "depth": args.depth},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": arm.describe(f", {steps} timed steps")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from colbert_b200 import _lib, kernels
    from colbert_b200.ranking import ColbertRanker

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    assert lib.cbk_device_supported(local_rank) == 1, "bench.py needs an sm_100 device"
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    dim, q_len = args.dim, args.q_len

    # Weak scaling: every GPU holds a shard of `--docs` documents (pids [rank*docs, (rank+1)*docs)); the job
    # scores `--queries * world` queries, each with `--cands` candidates drawn over the WHOLE corpus, so every
    # GPU scores ~queries*cands candidates per step.  All ranks see the same (replicated) queries and lists.
    store, doclens = build_store(torch, dev, args.docs, dim, dtype, seed=1234 + rank, doclen_fixed=args.doclen_fixed)
    ranker = ColbertRanker.from_store(store, doclens)
    if args.kernel == "tcgen05":
        ranker.kernel_flags |= _lib.CBK_FLAG_RERANK_TCGEN05
    elif args.kernel == "mma":
        ranker.kernel_flags &= ~_lib.CBK_FLAG_RERANK_TCGEN05
    n_queries = args.queries * world
    Q_host, cand_host = build_queries(torch, n_queries, q_len, dim, args.docs * world, args.cands, seed=4321)
    Q_pin, cand_pin = Q_host.pin_memory(), cand_host.pin_memory()
    Q_dev, cand_dev = Q_host.to(dev), cand_host.to(dev).reshape(-1).contiguous()
    n_cand_total = n_queries * args.cands
    rowptr = torch.arange(0, n_cand_total + 1, args.cands, dtype=torch.int64, device=dev)
    k = min(args.depth, args.cands)
    sharded = None
    if world > 1:
        from colbert_b200.sharding import ShardedColbertRanker
        from colbert_b200.ranking.colbert_ranker import torch_percentile
        all_dl = [torch.empty_like(doclens, device=dev) for _ in range(world)]
        dist.all_gather(all_dl, doclens.to(dev))
        gdl = torch.cat(all_dl).cpu()
        gstrides = sorted({torch_percentile(gdl, p) for p in (25, 50, 75)} | {int(gdl.max())})
        sharded = ShardedColbertRanker(ranker, rank * args.docs, gstrides)
    local = cand_dev - rank * args.docs
    mine = (local >= 0) & (local < args.docs)
    algo_bytes = int(ranker._doclens_dev[local[mine]].to(torch.int64).sum().item()) * dim * 2   # Σ doclen · dim · 2 B
    n_local_cands = int(mine.sum().item())
    del local, mine
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps: `value` -------------------------------------------------------------
    ev_k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def step_device(i=None):
        if sharded is None:
            pids_i, rowptr_i = cand_dev, rowptr
        else:                                            # route this shard's candidates (stays on the device)
            pids_i, rowptr_i = kernels.partition_candidates(cand_dev, rowptr, rank * args.docs, (rank + 1) * args.docs)
        if i is not None:
            ev_k0[i].record()
        scores = ranker.score_candidates(Q_dev, pids_i, rowptr_i)
        if i is not None:
            ev_k1[i].record()
        if sharded is None:
            return kernels.topk_per_query(scores, pids_i, rowptr_i, k, args.cands)
        keys = kernels.topk_per_query(scores, pids_i, rowptr_i, k, args.cands,
                                      flags=_lib.CBK_TOPK_NEG_INF_IS_PADDING, as_keys=True)
        return sharded._merge(sharded._exchange(keys), k)          # NCCL all-gather of packed keys + merge

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        out = step_device(i)
    t1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)) / args.steps

    # ---- where a sharded step spends its time (outside the timed region; reported as `phases_ms`) -------
    phases = None
    if sharded is not None:
        names = ["partition", "maxsim", "topk_keys", "all_gather", "merge"]
        acc = [0.0] * len(names)
        n_probe = 3
        for _ in range(n_probe):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            pids_i, rowptr_i = kernels.partition_candidates(cand_dev, rowptr, rank * args.docs, (rank + 1) * args.docs)
            ev[1].record()
            scores = ranker.score_candidates(Q_dev, pids_i, rowptr_i)
            ev[2].record()
            keys = kernels.topk_per_query(scores, pids_i, rowptr_i, k, args.cands, flags=_lib.CBK_TOPK_NEG_INF_IS_PADDING,
                                          as_keys=True)
            ev[3].record()
            gathered = sharded._exchange(keys)
            ev[4].record()
            sharded._merge(gathered, k)
            ev[5].record()
            torch.cuda.synchronize()
            for j in range(len(names)):
                acc[j] += ev[j].elapsed_time(ev[j + 1]) / n_probe
        pt = torch.tensor(acc, dtype=torch.float64, device=dev)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        phases = {n: round(v, 4) for n, v in zip(names, pt.tolist())}
        barrier()

    # ---- end to end through the public API with host buffers: `e2e` ---------------------------------
    # Public serving API: RerankPipeline.submit(pinned host inputs) / .result() → pinned host outputs.  Every step
    # copies its own inputs host→device and its results device→host inside the timed region; consecutive steps
    # are pipelined over two streams (the copies of step i+1 overlap the scoring of step i).
    from colbert_b200.ranking.pipeline import RerankPipeline
    pipe = RerankPipeline(sharded or ranker, n_queries, q_len, args.cands, depth=args.depth)

    def run_e2e(n_steps):
        prev, res = None, None
        for _ in range(n_steps):
            h = pipe.submit(Q_pin, cand_pin)
            if prev is not None:
                res = pipe.result(prev)             # host read of the previous step's result
            prev = h
        return pipe.result(prev)

    run_e2e(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = run_e2e(args.steps)
    e1.record()
    barrier()
    e2e_ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()

    times = torch.tensor([ms_total, e2e_ms_total, kern_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(algo_bytes), float(n_local_cands), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms_total, kern_ms = times.tolist()
    algo_bytes_all, scored_all, launches_all = sums.tolist()
    total_cands = n_cand_total                      # every candidate of every query is scored by exactly one GPU
    assert int(scored_all) == total_cands, (scored_all, total_cands)
    algo_bytes = algo_bytes_all / world             # per-GPU (per-launch) average for the roofline line
    launches = int(launches_all)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total_cands * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"rerank: {n_queries} queries x {args.cands} candidates ({args.queries}x{args.cands} per GPU), "
                            f"q_len {q_len}, dim {dim}, doclen {args.doclen_fixed or 'U[1,180]'}, {args.dtype} store of "
                            f"{args.docs} docs per GPU, top-{k} per query "
                            + (f"(BASELINE.json configs[{2 if args.doclen_fixed else 1}])" if dim == 128 else "(configs[1] at the author's width)"),
                "store_bytes_per_gpu": int(store.numel() * 2),
                "l2_policy": "inputs larger than L2: each step gathers "
                             f"{algo_bytes / 1e9:.1f} GB of distinct document rows from a {store.numel() * 2 / 1e9:.1f} GB store",
                "compute": ("16-bit store rows multiplied on tensor cores (m16n8k16, fp32 accumulate); a bf16 store is "
                            "converted to fp16 in registers so the query keeps 11 significant bits"),
                "parallelism": (f"store sharded by pid range over {world} GPUs, queries replicated, one NCCL all-gather "
                                f"of packed top-{k} keys per step + replicated merge") if world > 1 else "single GPU",
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (NCU_TRAFFIC_DEFAULT_BYTES if (world == 1 and args.queries == 4096 and args.cands == 1000
                                                                     and args.docs == 2_000_000 and not args.doclen_fixed
                                                                     and args.q_len == 32 and dim == 128) else None),
                         "kernel": "maxsim_rerank_kernel" if dim == 128 else "maxsim_wide_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                         "kernel_ms": kern_ms, "peak_source": peak_src},
            "e2e": {"value": total_cands * args.steps / (e2e_ms_total * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(pipe.h2d_bytes_per_step), "d2h_bytes_per_step": int(pipe.d2h_bytes_per_step),
                    "ms_per_step": e2e_ms_total / args.steps,
                    "api": "colbert_b200.ranking.pipeline.RerankPipeline.submit/result (pinned host in, pinned host out, "
                           "2-slot stream pipeline)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if phases is not None:
            line["phases_ms"] = phases          # max over ranks, measured on 3 extra steps after the timed region
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm(args.cpu_queries, args.cands, args.depth)
            best = min(arm.run() for _ in range(2))
            line["cpu_baseline"] = {"value": args.cpu_queries * args.cands / best, "unit": UNIT, "cores": arm.cores,
                                    "kind": "port", "sample": arm.describe(", best of 2 passes")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: spawn it ourselves so that `python bench.py --gpus N` works too
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
