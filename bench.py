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

# dram__bytes_read.sum + dram__bytes_write.sum of ONE maxsim_rerank_kernel launch of the default workload, from an
# `ncu --set full` capture (94.950 GB read + 0.019 GB written; algorithmic 94.869 GB — the difference is
# candidate/query metadata and DRAM sector granularity).  It is NOT measured by this run (a run under ncu is never a
# bench value): the line carries it with `traffic_source` naming the capture.  Only valid for the default arguments
# (same seeds → same candidate lists).
NCU_TRAFFIC_DEFAULT_BYTES = 94_968_766_096
NCU_TRAFFIC_SOURCE = "ncu --set full capture profiles/r01_end_rerank_ncu_full_summary.csv (not measured in this run)"


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` block (configs[2], [3]/[4])")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate")
    ap.add_argument("--e2e-exchange", default="p2p", choices=["allgather", "replicate", "p2p"],
                    help="N > 1: how RerankPipeline distributes the replicated inputs (see its docstring)")
    ap.add_argument("--wire", default="compact", choices=["compact", "native"],
                    help="what the serving entry is handed: compact = 16-bit queries / 32-bit pids (widened on the device), "
                         "native = the kernel's own fp32 / int64 (twice the bytes, no widening kernels)")
    ap.add_argument("--chunks", type=int, default=0, help="N > 1: query chunks per sharded step (0 = the ranker's default)")
    ap.add_argument("--parity-queries", type=int, default=8, help="queries re-scored by the oracle (parity gate)")
    ap.add_argument("--exh-docs", type=int, default=1_100_000, help="documents per GPU in the exhaustive secondary "
                    "(configs[3]: 8.8 M passages over 8 GPUs)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe): ONE nvidia-smi process for the whole run, every sample stamped with the
# host time it arrived; each timed region reports the samples that fall inside its own window
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        """Summary of the samples that arrived in [t0, t1] (host clock)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, power, reasons = [], [], [], set()
        for t, r in list(self.rows):
            if t < t0 or t > t1:
                continue
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
            try:
                power.append(float(parts[6]))
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "power_w_max": max(power) if power else None}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()


# --------------------------------------------------------------------------------------------------
# synthetic workload: unit-norm rows, doclen U[1,180], candidates uniform w/o replacement
# --------------------------------------------------------------------------------------------------
def build_store(torch, dev, n_docs, dim, dtype, seed, doclen_fixed=0, lo=1, hi=180):
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    if doclen_fixed:
        doclens = torch.full((n_docs,), int(doclen_fixed), dtype=torch.int64)
    else:
        doclens = torch.randint(lo, hi + 1, (n_docs,), generator=g, dtype=torch.int64)
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
# CPU arm: the reference's ranking path on the host cores — the reference's OWN code when oracle/_ref is staged
# (kind "reference": unmodified colbert/ranking/colbert_ranker.py + colbert/modeling/BaseModel.py imported through
# oracle/ref_loader.py), else the op-for-op torch port oracle/ref_port_torch.py (kind "port")
# --------------------------------------------------------------------------------------------------
class CpuArm:
    """rank_forward on a bounded sample: a 50k-doc store (configs[0] shape), queries of 32x128 fp32 with `n_cand`
    candidates each, all host threads."""

    def __init__(self, n_queries, n_cand, depth, seed=99, native_gpu=False):
        import tempfile
        import torch
        from colbert_b200 import synthetic
        from oracle import ref_loader
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        index = synthetic.make_index(seed, 50_000, dim=128, lo=1, hi=180)
        self.Q = torch.from_numpy(synthetic.make_queries(seed + 1, n_queries, 32, 128))
        cand = synthetic.make_candidates(seed + 2, n_queries, index.num_docs, n_cand)
        self.cand_lists = [c.tolist() for c in cand]
        self.n_queries, self.n_cand, self.depth, self.tokens = n_queries, n_cand, depth, index.num_tokens
        self.native_gpu = native_gpu
        if ref_loader.available():
            self.kind = "reference"
            cr, BaseModel, shims = ref_loader.load(cpu=not native_gpu)
            self.shims = shims
            import contextlib
            with tempfile.TemporaryDirectory() as d, shims(), contextlib.redirect_stdout(sys.stderr):
                synthetic.write_index(index, d)                     # (the reference prints its progress to stdout;
                self.ranker = cr.ColbertRanker(d, model=BaseModel, dim=128)   # stdout carries the ONE JSON line)
            self.what = ("the reference's own ColbertRanker.rank_forward + BaseModel.score (unmodified files staged in "
                         "oracle/_ref, " + ("DEVICE='cuda' as its author deploys it: CPU index_select -> pinned buffer "
                                            "-> H2D -> einsum on the GPU" if native_gpu else "DEVICE='cpu'") + ")")
        else:
            assert not native_gpu
            from oracle.ref_port_torch import CpuRankerPort
            self.kind = "port"
            store = torch.zeros(index.num_tokens + 512, 128, dtype=torch.float16)
            store[: index.num_tokens] = torch.from_numpy(index.emb)
            self.ranker = CpuRankerPort(store, index.doclens.tolist(), max_candidates=n_cand)
            self.shims = None
            self.what = "oracle/ref_port_torch.py (the reference's op sequence on torch CPU ops)"
        self.run(only=[0])                                              # warm-up

    def _call(self, b):
        Qt = self.Q[b].unsqueeze(0).permute(0, 2, 1)
        return self.ranker.rank_forward(Qt, self.cand_lists[b], depth=self.depth)

    def run(self, only=None):
        """one pass over the sample → seconds"""
        ctx = self.shims() if self.shims is not None else None
        if ctx is not None:
            ctx.__enter__()
        try:
            t0 = time.perf_counter()
            for b in (range(self.n_queries) if only is None else only):
                self._call(b)
            if self.native_gpu:
                self.torch.cuda.synchronize()
            return time.perf_counter() - t0
        finally:
            if ctx is not None:
                ctx.__exit__(None, None, None)

    def describe(self, extra=""):
        return (f"{self.n_queries} queries x {self.n_cand} candidates, Q 32x128 fp32, doclen U[1,180], fp16 store of "
                f"50k docs ({self.tokens} tokens), top-{self.depth}; {self.what}, {self.cores} threads{extra}")


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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.describe(f", {steps} timed steps")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # for context only (not the ratio's denominator): the reference the way its author deploys it, with a GPU for the
    # einsum — CPU gather into pinned staging buffers, PCIe, torch ops on the device, one query at a time
    try:
        import torch
        from oracle import ref_loader
        if ref_loader.available() and torch.cuda.is_available():
            hybrid = CpuArm(per_step_q, args.cands, args.depth, native_gpu=True)
            hybrid.run()
            best = min(hybrid.run() for _ in range(3))
            line["reference_native_gpu"] = {"value": per_step_q * args.cands / best, "unit": UNIT,
                                            "what": hybrid.describe(", best of 3 passes")}
    except Exception as exc:                                  # context only: never fail the arm on it
        line["reference_native_gpu"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# parity gate (outside every timed region): the oracle re-scores sampled queries from the very rows the GPU holds
# --------------------------------------------------------------------------------------------------
def fetch_rows(torch, dist, local_ranker, pids_np, rank, world):
    """Rows of the documents `pids_np` (GLOBAL pids, any order, any owner) → on rank 0
    ``(doclens int64 [n], rows fp32 [sum doclens, dim])`` in the order of `pids_np`; every rank contributes the
    documents of its own shard, read back from ITS OWN HBM store (what the kernels really scored)."""
    import numpy as np
    dev = local_ranker.device
    base, n_local = int(local_ranker.pid_base), int(local_ranker._doclens_dev.numel())
    mask = (pids_np >= base) & (pids_np < base + n_local)
    pos = np.nonzero(mask)[0]
    local = torch.from_numpy(pids_np[pos] - base).to(dev)
    lens = local_ranker._doclens_dev[local].to(torch.int64)
    starts = local_ranker._pfxsum_dev[local]
    total = int(lens.sum().item())
    excl = torch.cumsum(lens, 0) - lens
    idx = torch.arange(total, device=dev) - torch.repeat_interleave(excl, lens) + torch.repeat_interleave(starts, lens)
    rows = local_ranker.tensor[idx].float().cpu().numpy()
    payload = (pos, lens.cpu().numpy(), rows)
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
    else:
        gathered = [payload]
    if rank != 0:
        return None, None
    n = pids_np.shape[0]
    lens_all = np.zeros(n, dtype=np.int64)
    seen = np.zeros(n, dtype=np.int64)
    for pos_r, lens_r, _ in gathered:
        lens_all[pos_r] = lens_r
        seen[pos_r] += 1
    assert (seen == 1).all(), "every sampled document must be owned by exactly one rank"
    pf = np.concatenate([[0], np.cumsum(lens_all)])
    out = np.empty((int(pf[-1]), rows.shape[1]), dtype=np.float32)
    for pos_r, lens_r, rows_r in gathered:
        src = np.concatenate([[0], np.cumsum(lens_r)])
        for j, pcand in enumerate(pos_r):
            out[pf[pcand]: pf[pcand + 1]] = rows_r[src[j]: src[j + 1]]
    return lens_all, out


def parity_gate_rerank(torch, dist, local_ranker, strides, Q_host, cand_host, out_pids, out_scores, k, n_sample, rank,
                       world):
    """≥ n_sample queries of the timed batch: oracle.maxsim_exact on the rows pulled back from HBM (each rank its own
    shard's) + topk_desc, compared with what the timed step returned (check_topk: scores within 1e-3 relative, pids
    identical except inside ties).  At N > 1 this checks the REAL partition → MaxSim → NCCL all-gather → merge path."""
    import numpy as np
    n_q, n_cand = cand_host.shape
    sample = sorted(set(np.linspace(0, n_q - 1, n_sample).astype(int).tolist()))
    pids_np = cand_host[sample].reshape(-1).numpy()
    lens_all, rows = fetch_rows(torch, dist, local_ranker, pids_np, rank, world)
    if rank != 0:
        return None
    from oracle import maxsim_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_utils import SCORE_RTOL, check_topk
    pf = np.concatenate([[0], np.cumsum(lens_all)])
    got_p, got_s = out_pids.cpu().numpy(), out_scores.cpu().numpy()
    worst = 0.0
    for i, q in enumerate(sample):
        ids = np.arange(i * n_cand, (i + 1) * n_cand)
        ref = O.maxsim_exact(rows, lens_all, pf, strides, Q_host[q].numpy(), ids)
        cands = cand_host[q].numpy()
        rp, rs = O.topk_desc(ref, cands, k)
        fp, fs = O.topk_desc(ref, cands, None)
        worst = max(worst, check_topk(got_p[q], got_s[q], rp, rs, SCORE_RTOL, fp, fs))
    return {"queries": len(sample), "candidates_per_query": int(n_cand), "worst_rel": worst, "tol_rel": SCORE_RTOL,
            "checker": "oracle.maxsim_oracle.maxsim_exact + topk_desc on rows read back from each rank's HBM store; "
                       "tests/parity_utils.check_topk", "path": "single GPU" if world == 1 else
            f"partition -> MaxSim -> NCCL all-gather of keys -> merge over {world} ranks"}


def parity_gate_exhaustive(torch, dist, local_ranker, strides, Q_host, out_pids, out_scores, n_docs_total, n_sample,
                           rank, world, seed=7):
    """Exhaustive top-k cannot be recomputed on the CPU at full size; what is checked for sampled queries: (1) every
    winner's score equals the oracle's score of that document (1e-3 relative), (2) the list is score-descending with
    the pid tie-break, (3) 2,000 random non-winners per query all score at or below the k-th winner (+ tolerance)."""
    import numpy as np
    B, k = out_pids.shape
    sample = sorted(set(np.linspace(0, B - 1, n_sample).astype(int).tolist()))
    rng = np.random.default_rng(seed)
    got_p, got_s = out_pids.cpu().numpy(), out_scores.cpu().numpy()
    n_rand = 2000
    pid_lists = []
    for q in sample:
        others = rng.integers(0, n_docs_total, size=n_rand)
        pid_lists.append(np.concatenate([got_p[q], others]))
    pids_np = np.concatenate(pid_lists).astype(np.int64)
    lens_all, rows = fetch_rows(torch, dist, local_ranker, pids_np, rank, world)
    if rank != 0:
        return None
    from oracle import maxsim_oracle as O
    pf = np.concatenate([[0], np.cumsum(lens_all)])
    worst, per = 0.0, k + n_rand
    for i, q in enumerate(sample):
        ids = np.arange(i * per, (i + 1) * per)
        ref = O.maxsim_exact(rows, lens_all, pf, strides, Q_host[q].numpy(), ids)
        win, oth = ref[:k], ref[k:]
        rel = np.abs(got_s[q] - win) / np.maximum(np.abs(win), 1.0)
        worst = max(worst, float(rel.max()))
        assert rel.max() <= 1e-3, f"exhaustive parity: query {q} winner score error {rel.max()}"
        assert np.all(got_s[q][:-1] >= got_s[q][1:]), "winners are not score-descending"
        kth = float(got_s[q][-1])
        in_top = np.isin(pid_lists[i][k:], got_p[q])
        assert np.all(oth[~in_top] <= kth + 1e-3 * max(abs(kth), 1.0)), "a non-winner outscores the k-th winner"
    return {"queries": len(sample), "winners_per_query": int(k), "nonwinners_checked_per_query": n_rand,
            "worst_rel": worst, "tol_rel": 1e-3}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "tflops_burst": float(p["bf16_tflops"]), "source": "MEASURED_PEAKS.json (measured copy / cuBLAS bf16)"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0, "source": "B200_PROFILING.md fallback"}


def timed_region(torch, dist, world, step, warmup, min_steps, min_ms=1500.0, max_steps=400):
    """warm-up, then enough steps for the region to last ≥ min_ms (so that the clock sampler sees it under load) →
    (steps, total ms max over ranks, mean ms between the two events `step` records max over ranks, host window)."""
    dev = torch.cuda.current_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(None)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(None); b.record(); torch.cuda.synchronize()
    est = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    steps = int(min(max_steps, max(min_steps, -(-min_ms // max(float(est.item()), 1e-3)))))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    w0 = time.time()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        step(evs[i])
    t1.record()
    barrier()
    w1 = time.time()
    t = torch.tensor([t0.elapsed_time(t1), sum(x.elapsed_time(y) for x, y in evs) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return steps, float(t[0]), float(t[1]), (w0, w1)


def secondary_multiview(torch, sampler, dev, args, cand_dev, rowptr, d_view, q_view, dtype, peaks, dim=128, n_docs=None):
    """configs[2]: multi-view rerank (enable_multiview: every document is d_view rows, every query q_view rows, max over
    views = max over the document's rows), same 4096 x 1000 candidate lists as the headline.  d_view == 0: the headline
    workload itself (doclen U[1,180], 32 query rows) on a store of another dtype."""
    from colbert_b200 import kernels
    from colbert_b200.ranking import ColbertRanker
    n_docs = n_docs or args.docs
    store, doclens = build_store(torch, dev, n_docs, dim, dtype, seed=777 + d_view + dim, doclen_fixed=d_view)
    ranker = ColbertRanker.from_store(store, doclens)
    g = torch.Generator(device="cpu"); g.manual_seed(555 + q_view)
    n_q = rowptr.numel() - 1
    if n_docs != args.docs:                 # a smaller store: fold the candidate ids into it (lists stay duplicate-free enough)
        cand_dev = cand_dev % n_docs
    Q = torch.nn.functional.normalize(torch.randn(n_q, q_view, dim, generator=g), p=2, dim=2).to(dev)
    k = min(args.depth, args.cands)

    def step(ev):
        if ev: ev[0].record()
        scores = ranker.score_candidates(Q, cand_dev, rowptr)
        if ev: ev[1].record()
        return kernels.topk_per_query(scores, cand_dev, rowptr, k, args.cands)

    steps, ms_total, kern_ms, win = timed_region(torch, None, 1, step, args.warmup, args.steps)
    n_cand = cand_dev.numel()
    algo = int(ranker._doclens_dev[cand_dev].to(torch.int64).sum().item()) * dim * 2
    achieved = algo / (kern_ms * 1e-3) / 1e9
    dname = str(dtype).replace("torch.", "").replace("bfloat16", "bf16").replace("float16", "fp16")
    if d_view:
        what = (f"multi-view rerank: {n_q} queries x {args.cands} candidates, q_view {q_view}, d_view {d_view}, dim {dim}, "
                f"{dname} store of {n_docs} docs ({store.numel() * 2 / 1e9:.1f} GB), top-{k} (BASELINE.json configs[2]"
                + ("" if d_view == 8 else "; the author's dense.yaml operating point is 16 x 16")
                + (" at its own width 768, no projection" if dim == 768 else "") + ")")
    else:
        what = (f"rerank: {n_q} queries x {args.cands} candidates, q_len {q_view}, dim {dim}, doclen U[1,180], {dname} store of "
                f"{n_docs} docs ({store.numel() * 2 / 1e9:.1f} GB), top-{k} "
                + ("(configs[1] on the reference's own index dtype: the store whose scores meet 1e-3 against the reference's "
                   "goldens)" if dim == 128 else "(configs[1] at the author's un-projected width 768)"))
    out = {"workload": what,
           "value": n_cand * steps / (ms_total * 1e-3), "unit": UNIT, "steps": steps, "ms_per_step": ms_total / steps,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo,
                        "kernel": (("maxsim_wide_stream_kernel (tcgen05 streaming, fixed 16-row documents: no metadata lookups)"
                                    if d_view else "maxsim_wide_stream_kernel (tcgen05 streaming, ragged documents)") if dim != 128 else
                                   "maxsim_rerank_kernel" + (" (multi-view instantiation)" if d_view else "")), "traffic": None},
           "clocks": sampler.window(*win)}
    del store, ranker
    torch.cuda.empty_cache()
    return out


def secondary_exhaustive(torch, dist, sampler, dev, args, rank, world, peaks):
    """configs[3] / [4]: every document of this GPU's shard (configs[3]: 8.8 M passages over 8 GPUs = 1.1 M per GPU,
    doclen U[20,120] ~ 70 tokens, 128-d fp16, 19.7 GB) against Nq queries, top-1000 per query; at N > 1 the sharded
    rank_exhaustive: local top-1000 as packed keys -> NCCL all-gather -> replicated merge."""
    from colbert_b200 import kernels
    from colbert_b200.ranking import ColbertRanker
    from colbert_b200.ranking.colbert_ranker import torch_percentile
    from colbert_b200.sharding import ShardedColbertRanker
    store, doclens = build_store(torch, dev, args.exh_docs, 128, torch.float16, seed=4242 + rank, lo=20, hi=120)
    ranker = ColbertRanker.from_store(store, doclens)
    sharded = None
    if world > 1:
        all_dl = [torch.empty_like(doclens, device=dev) for _ in range(world)]
        dist.all_gather(all_dl, doclens.to(dev))
        gdl = torch.cat(all_dl).cpu()
        gstrides = sorted({torch_percentile(gdl, p) for p in (25, 50, 75)} | {int(gdl.max())})
        sharded = ShardedColbertRanker(ranker, rank * args.exh_docs, gstrides)
    rows = int(ranker.doclens_pfxsum[-1])
    k = min(1000, args.exh_docs)
    out = {"workload": f"exhaustive MaxSim: {args.exh_docs} documents per GPU x {world} GPU(s) = "
                       f"{args.exh_docs * world / 1e6:.1f} M passages, doclen U[20,120], dim 128, fp16 store "
                       f"({rows * 256 / 1e9:.1f} GB per GPU), q_len 32, top-{k} per query"
                       + (", NCCL all-gather of packed keys + replicated merge" if world > 1 else "")
                       + " (BASELINE.json configs[3] shard shape; Nq sweep = configs[4])", "unit": "documents scored/s"}
    for nq in (1, 16, 64):
        g = torch.Generator(device="cpu"); g.manual_seed(9000 + nq)
        Q_host = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g), p=2, dim=2)
        Q = Q_host.to(dev)
        result = {}

        def step(ev):
            if ev: ev[0].record()
            dense = ranker.score_all(Q)
            if ev: ev[1].record()
            if sharded is None:
                s, p = kernels.topk_dense(dense, k, pid_base=0)
            else:
                keys = kernels.topk_dense(dense, k, pid_base=sharded.pid_base, as_keys=True)
                p, s = sharded._merge(sharded._exchange(keys), k)
            result["p"], result["s"] = p, s

        steps, ms_total, kern_ms, win = timed_region(torch, dist, world, step, args.warmup, args.steps)
        flops = 2.0 * 32 * nq * 128 * rows
        entry = {"value": nq * args.exh_docs * world * steps / (ms_total * 1e-3), "steps": steps,
                 "ms_per_step": ms_total / steps, "maxsim_kernel_ms": kern_ms, "clocks": sampler.window(*win)}
        if nq == 1:
            ach = rows * 256 / (kern_ms * 1e-3) / 1e9
            entry["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": rows * 256}
        else:
            ach = flops / (kern_ms * 1e-3) / 1e12
            entry["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": ach / peaks["tflops_sustained"], "frac_of_burst_peak": ach / peaks["tflops_burst"],
                                 "algorithmic_flops_per_launch": flops}
        entry["roofline"]["kernel"] = "maxsim_exhaustive_kernel (tcgen05)"
        if nq == 16 and not args.no_parity:
            entry["parity"] = parity_gate_exhaustive(torch, dist, ranker, sharded.strides if sharded else ranker.strides,
                                                     Q_host, result["p"], result["s"], args.exh_docs * world, 2, rank, world)
        out[f"Nq{nq}"] = entry
    del store, ranker, sharded
    torch.cuda.empty_cache()
    return out


def secondary_allpairs_training(torch, sampler, dev, args, peaks):
    """BaseModel.score as the reference TRAINS with it — forward + backward over the all-gathered batch
    of colbert_model.py:87-95 at the author's width (q = 170 questions x 32 rows, d = 340 passages x 384 rows, h = 768)."""
    from colbert_b200 import kernels
    from colbert_b200.modeling.BaseModel import BaseModel
    nq, m, nd, n, h = 170, 32, 340, 384, 768
    g = torch.Generator(device=dev); g.manual_seed(31337)
    Q = torch.nn.functional.normalize(torch.randn(nq, m, h, generator=g, device=dev), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(nd, n, h, generator=g, device=dev), dim=-1)
    qmask = torch.ones(nq, m, dtype=torch.int64, device=dev)
    dmask = (torch.arange(n, device=dev)[None, :] < torch.randint(n // 4, n + 1, (nd, 1), generator=g, device=dev)).long()
    W = torch.randn(nq, nd, generator=g, device=dev)
    Qp = kernels.mask_cast_rows(Q.reshape(-1, h), qmask.reshape(-1), torch.float16).reshape(nq, m, h)
    Dp = kernels.mask_cast_rows(D.reshape(-1, h), dmask.reshape(-1), torch.float16).reshape(nd, n, h)
    flops = 2.0 * nq * m * nd * n * h
    res = {}

    def fwd_only(ev):
        if ev: ev[0].record()
        res["s"], res["a"] = kernels.score_allpairs_fwd(Qp, Dp)
        if ev: ev[1].record()

    steps, ms_total, fwd_ms, win = timed_region(torch, None, 1, fwd_only, args.warmup, args.steps, min_ms=1000.0)
    clocks_fwd = sampler.window(*win)

    def train_step(ev):
        Qa, Da = Q.detach().requires_grad_(True), D.detach().requires_grad_(True)
        if ev: ev[0].record()
        (BaseModel.score(Qa, Da, qmask, dmask) * W).sum().backward()
        if ev: ev[1].record()
        res["dQ"], res["dD"] = Qa.grad, Da.grad

    steps2, ms_total2, step_ms, win2 = timed_region(torch, None, 1, train_step, args.warmup, args.steps, min_ms=1000.0)
    ach = flops / (fwd_ms * 1e-3) / 1e12
    out = {"workload": f"all-pairs MaxSim with backward (training shape, colbert_model.py:87-95): q {nq} x m {m}, d {nd} x n {n}, "
                       f"h {h}; 16-bit operands (fp16), fp32 accumulation; inputs 0.43 GB fp32 (> L2)",
           "forward_kernel_ms": fwd_ms, "steps": steps,
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tflops_sustained"], "frac_of_burst_peak": ach / peaks["tflops_burst"],
                        "algorithmic_flops_per_launch": flops, "kernel": "score_allpairs_fwd_kernel (tcgen05)"},
           "score_fwd_bwd_ms": step_ms, "score_fwd_bwd_steps": steps2,
           "score_fwd_bwd_what": "BaseModel.score(Q, D, q_mask, d_mask) under autograd: 2 mask+cast passes, forward kernel, "
                                 "dQ and dD kernels, fp32 gradients",
           "clocks": clocks_fwd, "clocks_fwd_bwd": sampler.window(*win2)}
    if not args.no_parity:
        # oracle (numpy restatement of BaseModel.py:41-45 and its autograd) on a slice: 3 queries x 5 documents
        import numpy as np
        from oracle import maxsim_oracle as O
        qs, ds = [0, nq // 2, nq - 1], [0, 1, nd // 2, nd - 2, nd - 1]
        Qh, Dh = Q[qs].half().float().cpu().numpy(), D[ds].half().float().cpu().numpy()
        rs, _, _, rarg = O.score_allpairs_grad(Qh, Dh, qmask[qs].cpu().numpy(), dmask[ds].cpu().numpy(),
                                               np.zeros((len(qs), len(ds)), np.float32))
        got = res["s"][qs][:, ds].cpu().numpy()
        worst = float(np.abs(got - rs).max() / max(1.0, np.abs(rs).max()))
        arg_ok = bool((res["a"][qs][:, ds].cpu().numpy() == rarg).all())
        # whole batch against the reference's op sequence in fp32 on the GPU (BaseModel.py:41-45 on the 16-bit-rounded
        # inputs): scores; the kernel's arg-max is a true arg-max (an fp32 matmul and the tensor cores sum in different
        # orders, so among ~2 M (q, d, m) maxima a few near-ties legitimately resolve differently — each is checked to be
        # within 1e-5 of the maximum); and dQ / dD equal what autograd derives for THAT arg-max (gather / index_add).
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            Qm_, Dm_ = Q.half().float() * qmask[..., None], D.half().float() * dmask[..., None]
            dQ_ref, dD_ref = torch.zeros_like(Qm_), torch.zeros_like(Dm_)
            doc_row0 = (torch.arange(nd, device=dev) * n)[None, :, None]
            s_worst = tie_gap = 0.0
            flips = 0
            for lo in range(0, nq, 10):
                hi = min(nq, lo + 10)
                sim = torch.einsum("qmh,dnh->qdmn", Qm_[lo:hi], Dm_)
                mx, ix = sim.max(-1)
                ours = res["a"][lo:hi].long()
                at = sim.gather(-1, ours.unsqueeze(-1)).squeeze(-1)
                tie_gap = max(tie_gap, float((mx - at).max()))
                flips += int((ix != ours).sum())
                ref_s = mx.sum(-1)
                s_worst = max(s_worst, float(((res["s"][lo:hi] - ref_s).abs() / ref_s.abs().clamp_min(1.0)).max()))
                del sim, mx, ix, at
                rows = (doc_row0 + ours).reshape(-1)                                        # [c * nd * m] rows of D
                picked = Dm_.reshape(nd * n, h)[rows].view(hi - lo, nd, m, h)
                dQ_ref[lo:hi] = (picked * W[lo:hi, :, None, None]).sum(1)
                del picked
                contrib = (W[lo:hi, :, None, None] * Qm_[lo:hi, None, :, :]).reshape(-1, h)
                dD_ref.view(nd * n, h).index_add_(0, rows, contrib)
                del contrib
            dQ_ref *= qmask[..., None]
            dD_ref *= dmask[..., None]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        gq = float((res["dQ"] - dQ_ref).abs().max() / dQ_ref.abs().max())
        gd = float((res["dD"] - dD_ref).abs().max() / dD_ref.abs().max())
        assert worst <= 1e-3 and arg_ok and s_worst <= 1e-3 and tie_gap <= 1e-5 and gq <= 1e-3 and gd <= 1e-3, \
            (worst, arg_ok, s_worst, tie_gap, gq, gd)
        out["parity"] = {"scores_worst_rel_vs_oracle": worst, "argmax_equal_to_oracle": arg_ok, "queries": len(qs), "documents": len(ds),
                         "scores_worst_rel_vs_torch_fp32_full_batch": s_worst,
                         "argmax_differs_from_torch_fp32": flips, "of": nq * nd * m, "largest_gap_to_the_maximum": tie_gap,
                         "grad_Q_worst_rel": gq, "grad_D_worst_rel": gd,
                         "grad_reference": "autograd's formula for the kernel's own arg-max (gather / index_add in torch fp32)",
                         "tol_rel": 1e-3}
        del dQ_ref, dD_ref, Qm_, Dm_
    torch.cuda.empty_cache()
    return out


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
    peaks = load_peaks()
    sampler = ClockSampler(local_rank).start()

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
    # what the serving entry is handed: the encoder's 16-bit query embeddings and 32-bit pids when the kernel rounds the
    # query to fp16 anyway (bit-identical scores, tests/test_gpu_parity.py::test_rerank_pipeline_matches_direct_call),
    # else fp32 / int64
    compact = args.wire == "compact"
    wire16 = compact and bool(ranker.query_rounded_to_fp16)
    Q_pin = (Q_host.to(torch.float16) if wire16 else Q_host).pin_memory()
    cand_pin = (cand_host.to(torch.int32) if compact and args.docs * world < 2 ** 31 else cand_host).pin_memory()
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

    cand_2d = cand_dev.view(n_queries, args.cands)
    if sharded is not None:
        sharded.maxsim_events = []

    def step_device(i=None):
        """→ (pids [B,k], scores [B,k])"""
        if sharded is None:
            if i is not None:
                ev_k0[i].record()
            scores = ranker.score_candidates(Q_dev, cand_dev, rowptr)
            if i is not None:
                ev_k1[i].record()
            s, p = kernels.topk_per_query(scores, cand_dev, rowptr, k, args.cands)
            return p, s
        # sharded: the ranker's own batched call — per query chunk: partition -> MaxSim -> local top-k keys, then
        # (second stream, under the next chunk's MaxSim) NCCL all-gather of packed keys + replicated merge
        return sharded.rank_forward_batch(Q_dev, cand_2d, depth=k, chunks=args.chunks or None)

    for _ in range(args.warmup):
        step_device()
    barrier()
    w0 = time.time()
    launches0 = _lib.launch_count()
    if sharded is not None:
        sharded.maxsim_events = []          # only the timed steps from here on
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        out_pids, out_scores = step_device(i)
    t1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    if sharded is None:
        kern_ms = sum(a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)) / args.steps
    else:       # the MaxSim launches of the timed steps (one per query chunk), summed per step
        kern_ms = sum(a.elapsed_time(b) for a, b in sharded.maxsim_events) / args.steps
        sharded.maxsim_events = None

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
    pipe = RerankPipeline(sharded or ranker, n_queries, q_len, args.cands, depth=args.depth, input_exchange=args.e2e_exchange)

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
    w1 = time.time()
    e2e_ms_total = e0.elapsed_time(e1)
    clocks = sampler.window(w0, w1)
    # the serving entry must return what the device-resident step returned (same inputs): every rank checks the
    # slice of the result it holds
    e2e_pids, e2e_scores = res
    lo_q, hi_q = pipe.result_rows
    e2e_same = bool(torch.equal(e2e_pids, out_pids[lo_q:hi_q].cpu()) and torch.equal(e2e_scores, out_scores[lo_q:hi_q].cpu()))

    # ---- parity gate: the oracle on the same inputs, in the same process, outside the timed regions ----------
    parity = None
    if not args.no_parity:
        parity = parity_gate_rerank(torch, dist, ranker, sharded.strides if sharded else ranker.strides, Q_host, cand_host,
                                    out_pids, out_scores, k, args.parity_queries, rank, world)

    times = torch.tensor([ms_total, e2e_ms_total, kern_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(algo_bytes), float(n_local_cands), float(launches), float(e2e_same)], dtype=torch.float64,
                        device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms_total, kern_ms = times.tolist()
    algo_bytes_all, scored_all, launches_all, same_all = sums.tolist()
    total_cands = n_cand_total                      # every candidate of every query is scored by exactly one GPU
    assert int(scored_all) == total_cands, (scored_all, total_cands)
    assert int(same_all) == world, "the serving entry returned something else than the device-resident step"
    algo_bytes = algo_bytes_all / world             # per-GPU (per-launch) average for the roofline line
    launches = int(launches_all)
    h2d, d2h, e2e_api = int(pipe.h2d_bytes_per_step), int(pipe.d2h_bytes_per_step), pipe.describe()
    store_bytes = int(store.numel() * 2)
    del pipe, store, ranker, sharded, Q_dev
    torch.cuda.empty_cache()

    # ---- secondary workloads (configs[2], [3], [4]); same process, own timed regions and clock windows ----------
    secondary = None
    default_shape = dim == 128 and not args.doclen_fixed and q_len == 32
    if not args.no_secondary and default_shape:
        secondary = {}
        if world == 1:
            secondary["multiview_8x8"] = secondary_multiview(torch, sampler, dev, args, cand_dev, rowptr, 8, 8, dtype, peaks)
            secondary["multiview_16x16"] = secondary_multiview(torch, sampler, dev, args, cand_dev, rowptr, 16, 16, dtype, peaks)
            secondary["multiview_16x16_dim768"] = secondary_multiview(torch, sampler, dev, args, cand_dev, rowptr, 16, 16, dtype,
                                                                      peaks, dim=768, n_docs=min(args.docs, 1_000_000))
            secondary["rerank_dim768"] = secondary_multiview(torch, sampler, dev, args, cand_dev[: min(n_queries, 1024) * args.cands],
                                                             rowptr[: min(n_queries, 1024) + 1], 0, 32, dtype, peaks, dim=768,
                                                             n_docs=min(args.docs, 300_000))
            if dtype != torch.float16:
                secondary["rerank_fp16_store"] = secondary_multiview(torch, sampler, dev, args, cand_dev, rowptr, 0, 32,
                                                                     torch.float16, peaks)
        del cand_dev
        torch.cuda.empty_cache()
        secondary["exhaustive"] = secondary_exhaustive(torch, dist, sampler, dev, args, rank, world, peaks)
        if world == 1:
            secondary["allpairs_training"] = secondary_allpairs_training(torch, sampler, dev, args, peaks)
    sampler.stop()

    if rank == 0:
        peak, peak_src = peaks["hbm_gbs"], peaks["source"]
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        default_traffic = (world == 1 and args.queries == 4096 and args.cands == 1000 and args.docs == 2_000_000
                           and default_shape)
        line = {
            "metric": METRIC, "value": total_cands * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"rerank: {n_queries} queries x {args.cands} candidates ({args.queries}x{args.cands} per GPU), "
                            f"q_len {q_len}, dim {dim}, doclen {args.doclen_fixed or 'U[1,180]'}, {args.dtype} store of "
                            f"{args.docs} docs per GPU, top-{k} per query "
                            + (f"(BASELINE.json configs[{2 if args.doclen_fixed else 1}])" if dim == 128 else "(configs[1] at the author's width)"),
                "store_bytes_per_gpu": store_bytes,
                "l2_policy": "inputs larger than L2: each step gathers "
                             f"{algo_bytes / 1e9:.1f} GB of distinct document rows from a {store_bytes / 1e9:.1f} GB store",
                "compute": ("16-bit store rows multiplied on tensor cores (m16n8k16, fp32 accumulate); a bf16 store is "
                            "converted to fp16 in registers so the query keeps 11 significant bits"),
                "parallelism": (f"store sharded by pid range over {world} GPUs, queries replicated, one NCCL all-gather "
                                f"of packed top-{k} keys per step + replicated merge") if world > 1 else "single GPU",
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_DEFAULT_BYTES if default_traffic else None,
                         "traffic_source": NCU_TRAFFIC_SOURCE if default_traffic else None,
                         "kernel": "maxsim_rerank_kernel" if dim == 128 else "maxsim_wide_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                         "kernel_ms": kern_ms, "peak_source": peak_src},
            "e2e": {"value": total_cands * args.steps / (e2e_ms_total * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_total / args.steps, "matches_device_step": True, "api": e2e_api},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if parity is not None:
            line["parity"] = parity
        if phases is not None:
            line["phases_ms"] = phases          # max over ranks, measured on 3 extra steps after the timed region
        if secondary is not None:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm(args.cpu_queries, args.cands, args.depth)
            best = min(arm.run() for _ in range(2))
            line["cpu_baseline"] = {"value": args.cpu_queries * args.cands / best, "unit": UNIT, "cores": arm.cores,
                                    "kind": arm.kind, "sample": arm.describe(", best of 2 passes")}
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
