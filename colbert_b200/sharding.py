"""Document-range sharding of the embedding store across the GPUs of one box.

The reference's ranker is single-GPU (`DEVICE = "cuda"`, colbert/ranking/colbert_ranker.py:12); this
is the multi-GPU layout of the same path: contiguous pid ranges balanced by TOKEN count (the store is
what fills HBM), one process per GPU, queries replicated, every rank scores the candidates that fall
in its range, and the per-rank top-k lists are exchanged with ONE all-gather of packed
(score, pid) keys and merged identically on every rank.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import kernels
from ._lib import CBK_FLAG_SKIP_FOREIGN_PIDS, CBK_TOPK_NEG_INF_IS_PADDING


def plan_shards(doclens_pfxsum: torch.Tensor, world: int) -> List[int]:
    """→ pid boundaries ``b[0]=0 ≤ b[1] ≤ … ≤ b[world]=N``; rank r owns pids ``[b[r], b[r+1])`` and store
    rows ``[pfxsum[b[r]], pfxsum[b[r+1]])``.  Boundaries are the pids whose prefix sum is closest to
    ``r/world`` of the token total, so shards differ by at most one document's worth of tokens."""
    assert doclens_pfxsum.dim() == 1 and doclens_pfxsum.numel() >= 1 and world >= 1
    n_docs = doclens_pfxsum.numel() - 1
    total = int(doclens_pfxsum[-1])
    bounds = [0]
    for r in range(1, world):
        target = (total * r) // world
        j = int(torch.searchsorted(doclens_pfxsum, torch.tensor(target, dtype=doclens_pfxsum.dtype)))
        j = min(max(j, 0), n_docs)
        if j > 0 and abs(int(doclens_pfxsum[j - 1]) - target) <= abs(int(doclens_pfxsum[j]) - target):
            j -= 1
        bounds.append(max(j, bounds[-1]))
    bounds.append(n_docs)
    return bounds


def owner_of(pids: torch.Tensor, bounds: List[int]) -> torch.Tensor:
    """Rank owning each global pid."""
    b = torch.as_tensor(bounds[1:-1], dtype=pids.dtype, device=pids.device)
    return torch.searchsorted(b, pids, right=True)


class ShardedColbertRanker:
    """One rank's view of a pid-range-sharded store.

    ``local`` is a :class:`colbert_b200.ranking.ColbertRanker` holding only this rank's documents;
    ``pid_base`` its first global pid; ``global_strides`` the strides of the WHOLE corpus (the
    reference's zero-floor rule, SURVEY.md §8 a12′, depends on them, so every shard must use the same
    list).  ``rank_forward_batch`` returns the same (pids, scores) on every rank, equal to what a
    single-GPU ranker over the whole corpus returns."""

    def __init__(self, local, pid_base: int, global_strides, group: Optional[dist.ProcessGroup] = None):
        self.local = local
        self.pid_base = int(pid_base)
        self.strides = [int(s) for s in global_strides]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if local is not None:
            local.strides = self.strides
            local.pid_base = self.pid_base
            local.kernel_flags |= CBK_FLAG_SKIP_FOREIGN_PIDS
        self._comm_stream = None
        #: when set to a list, every MaxSim launch of rank_forward_batch appends its (start, end) CUDA events to it
        self.maxsim_events = None

    @classmethod
    def from_global_tensors(cls, embeddings: torch.Tensor, doclens, device, group=None, store_dtype=None):
        """Every rank is handed the whole (host) corpus and keeps its own pid range — for tests and
        small indexes; large stores are loaded shard by shard instead."""
        from .ranking.colbert_ranker import ColbertRanker, torch_percentile
        doclens = torch.as_tensor(doclens, dtype=torch.int64)
        pfx = torch.zeros(doclens.numel() + 1, dtype=torch.int64)
        torch.cumsum(doclens, 0, out=pfx[1:])
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        bounds = plan_shards(pfx, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        strides = sorted({torch_percentile(doclens, p) for p in (25, 50, 75)} | {int(doclens.max())})
        local = ColbertRanker.from_tensors(embeddings[int(pfx[lo]): int(pfx[hi])], doclens[lo:hi].tolist(),
                                           device=device, store_dtype=store_dtype)
        self = cls(local, lo, strides, group)
        self.bounds = bounds
        return self

    # ---- the four stages; the CPU (gloo) tests override the two device stages -----------------------
    def _local_topk_keys(self, Q: torch.Tensor, cand_pids: torch.Tensor, cand_rowptr: torch.Tensor, k: int,
                         max_cand: int, q_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Route this shard's share of every candidate list, score it, keep the local top-k as packed keys."""
        n_docs = self.local.doclens.numel()
        my_pids, my_rowptr = kernels.partition_candidates(cand_pids, cand_rowptr, self.pid_base, self.pid_base + n_docs)
        if self.maxsim_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        scores = self.local.score_candidates(Q, my_pids, my_rowptr, q_lens)
        if self.maxsim_events is not None:
            ev[1].record()
            self.maxsim_events.append(ev)
        return kernels.topk_per_query(scores, my_pids, my_rowptr, k, max_cand,
                                      flags=CBK_TOPK_NEG_INF_IS_PADDING, as_keys=True)

    def _exchange(self, keys: torch.Tensor) -> torch.Tensor:
        """[B, k] → [world, B, k]: one all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
        if self.world == 1:
            return keys.unsqueeze(0)
        B, k = keys.shape
        gathered = torch.empty((self.world * B, k), dtype=keys.dtype, device=keys.device)   # rank-major concat
        dist.all_gather_into_tensor(gathered, keys.contiguous(), group=self.group)
        return gathered.view(self.world, B, k)

    def _merge(self, gathered: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        scores, pids = kernels.merge_topk_keys(gathered, k)
        return pids, scores

    def _local_exhaustive_keys(self, Q: torch.Tensor, k: int) -> torch.Tensor:
        """Every document of this shard against every query; local top-k as packed keys (global pids)."""
        return kernels.topk_dense(self.local.score_all(Q), k, pid_base=self.pid_base, as_keys=True)

    def _local_num_docs(self) -> int:
        return int(self.local.doclens.numel())

    def rank_exhaustive(self, Q: torch.Tensor, k: int = 1000):
        """Exhaustive scoring over the whole sharded corpus: every rank scans its own
        shard, one all-gather of [B, k] packed keys, replicated merge → (pids [B,k], scores [B,k]), identical on
        all ranks and equal to a single-GPU ``ColbertRanker.rank_exhaustive`` over the whole corpus."""
        dev = self.local.device if self.local is not None else Q.device
        Q = Q.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        k_local = min(int(k), self._local_num_docs())
        keys = self._local_exhaustive_keys(Q, k_local)
        if k_local < k:   # a shard smaller than k: pad its list so that every rank contributes [B, k]
            pad = torch.zeros((keys.size(0), k - k_local), dtype=keys.dtype, device=keys.device)
            keys = torch.cat([keys, pad], dim=1)
        return self._merge(self._exchange(keys), int(k))

    def rank_forward_batch(self, Q: torch.Tensor, cand_pids: torch.Tensor, cand_rowptr: Optional[torch.Tensor] = None,
                           depth: Optional[int] = 10, max_cand: Optional[int] = None, chunks: Optional[int] = None,
                           q_lens: Optional[torch.Tensor] = None):
        """Same contract as ``ColbertRanker.rank_forward_batch`` with GLOBAL pids; identical on all ranks.

        ``chunks`` (equal-length lists ``[B, n]`` on more than one GPU): the batch is scored in that many query chunks,
        and the key all-gather + merge of chunk i run on a second stream underneath the MaxSim of chunk i+1 — the
        exchange is latency- and skew-bound (it waits for the slowest rank).  Measured on 8 x B200 (4096 queries x 1000
        candidates per GPU, `bench.py --gpus 8 --chunks c`): 16.25 / 16.40 / 16.72 ms per step with 1 / 2 / 4 chunks — the
        persistent MaxSim kernel fills every SM, so the NCCL kernel of chunk i cannot start before chunk i + 1 drains and
        each chunk adds its own launch tail: the default is ONE chunk; the option stays for short-kernel regimes.
        The result does not depend on it."""
        dev = self.local.device if self.local is not None else Q.device
        Q = Q.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        B = Q.size(0)
        cand_pids = cand_pids.to(dev, non_blocking=True)
        extra = {}
        if q_lens is not None:
            extra["q_lens"] = torch.as_tensor(q_lens).to(dev, dtype=torch.int32, non_blocking=True).contiguous()
        dense = cand_rowptr is None
        if dense:
            assert cand_pids.dim() == 2 and cand_pids.size(0) == B
            n = cand_pids.size(1)
            max_cand = n
            if chunks is None:
                chunks = 1
            chunks = max(1, min(int(chunks), B))
            if chunks > 1 and Q.is_cuda:
                return self._rank_forward_chunked(Q, cand_pids.contiguous(), n, depth, chunks, extra.get("q_lens"))
            cand_rowptr = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=dev)
            cand_pids = cand_pids.reshape(-1)
        else:
            cand_rowptr = cand_rowptr.to(dev, non_blocking=True)
            if max_cand is None:
                max_cand = int((cand_rowptr[1:] - cand_rowptr[:-1]).max().item())
        cand_pids = cand_pids.contiguous()
        k = max_cand if depth is None else min(int(depth), max_cand)
        keys = self._local_topk_keys(Q, cand_pids, cand_rowptr, k, max_cand, **extra)
        return self._merge(self._exchange(keys), k)

    def _rank_forward_chunked(self, Q: torch.Tensor, cand: torch.Tensor, n: int, depth: Optional[int], chunks: int,
                              q_lens: Optional[torch.Tensor] = None):
        dev = Q.device
        B = Q.size(0)
        k = n if depth is None else min(int(depth), n)
        compute = torch.cuda.current_stream(dev)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        comm = self._comm_stream
        out_p = torch.empty((B, k), dtype=torch.int64, device=dev)
        out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        comm.wait_stream(compute)                                   # out_* exist before the side stream writes them
        for c in range(chunks):
            q0, q1 = B * c // chunks, B * (c + 1) // chunks
            rowptr = torch.arange(0, (q1 - q0 + 1) * n, n, dtype=torch.int64, device=dev)
            keys = self._local_topk_keys(Q[q0:q1], cand[q0:q1].reshape(-1), rowptr, k, n,
                                         **({} if q_lens is None else {"q_lens": q_lens[q0:q1].contiguous()}))
            ready = torch.cuda.Event()
            ready.record(compute)
            keys.record_stream(comm)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                p, s = self._merge(self._exchange(keys), k)
                out_p[q0:q1].copy_(p)
                out_s[q0:q1].copy_(s)
        compute.wait_stream(comm)
        out_p.record_stream(comm)
        out_s.record_stream(comm)
        return out_p, out_s
