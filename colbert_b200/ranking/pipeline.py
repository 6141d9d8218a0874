"""Batched serving entry for the rerank path.

The reference serves queries one at a time in a Python loop (``DenseRetrieverServer.retrieve``,
colbert/training/dense_server_client.py:36-49: encode → per query ``ColbertRetriever.search`` →
``rank_forward`` with a blocking ``.cpu()`` per stride bucket).  Here a step is a whole batch of queries
with their candidate lists, and consecutive steps are pipelined over CUDA streams:

    copy stream     H2D(step i+1)                    H2D(step i+2)
    compute stream  ............  MaxSim+top-k(i) → D2H(i)  MaxSim+top-k(i+1) → D2H(i+1)

so the host↔device copies of a step hide behind the scoring of its neighbour.  With a sharded store
(``ShardedColbertRanker``) every rank uploads only its 1/world slice of the replicated inputs and the ranks
all-gather the slices over NVLink instead of each pulling the whole batch through its own PCIe link.

Inputs must be PINNED host tensors (``tensor.pin_memory()``); results come back in pinned host buffers owned
by the pipeline (valid until the slot is reused two submissions later).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


class RerankPipeline:
    def __init__(self, ranker, n_queries: int, q_len: int, n_cand: int, depth: int = 10, slots: int = 2):
        """``ranker``: ``ColbertRanker`` or ``ShardedColbertRanker``.  Every step carries ``n_queries`` queries
        of ``q_len`` rows and ``n_cand`` candidates each (equal-length lists)."""
        self.ranker = ranker
        local = getattr(ranker, "local", ranker)
        self.device = local.device
        self.dim = local.dim
        sharded = hasattr(ranker, "local")                      # ShardedColbertRanker (ColbertRanker.rank is a method)
        self.world = ranker.world if sharded else 1
        self.rank = ranker.rank if sharded else 0
        self.group = ranker.group if sharded else None
        self.n_queries, self.q_len, self.n_cand = n_queries, q_len, n_cand
        self.k = min(int(depth), n_cand)
        self.depth = depth
        self.slots = slots
        assert n_queries % self.world == 0, "the batch must split evenly over the ranks"
        self.slice = n_queries // self.world
        dev = self.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.Q_dev = [torch.empty((n_queries, q_len, self.dim), dtype=torch.float32, device=dev) for _ in range(slots)]
        self.C_dev = [torch.empty((n_queries, n_cand), dtype=torch.int64, device=dev) for _ in range(slots)]
        self.out_pids = [torch.empty((n_queries, self.k), dtype=torch.int64).pin_memory() for _ in range(slots)]
        self.out_scores = [torch.empty((n_queries, self.k), dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_done = [torch.cuda.Event() for _ in range(slots)]
        self._i = 0
        # the input all-gather runs on the copy stream, concurrently with the key all-gather of the previous step on
        # the compute stream: it needs its own communicator (collective call: every rank builds its pipeline)
        self.in_group = dist.new_group(ranks=list(range(self.world))) if self.world > 1 else None
        self.h2d_bytes_per_step = (self.slice * q_len * self.dim * 4 + self.slice * n_cand * 8) * self.world
        self.d2h_bytes_per_step = n_queries * self.k * 12 * self.world

    def submit(self, Q_host: torch.Tensor, cand_host: torch.Tensor) -> int:
        """Queue one step; returns the slot to pass to :meth:`result`.  ``Q_host`` ``[n_queries, q_len, dim]`` fp32
        and ``cand_host`` ``[n_queries, n_cand]`` int64, both pinned and identical on every rank."""
        assert Q_host.is_pinned() and cand_host.is_pinned(), "inputs must be pinned host tensors"
        s = self._i % self.slots
        self._i += 1
        compute = torch.cuda.current_stream(self.device)
        lo, hi = self.rank * self.slice, (self.rank + 1) * self.slice
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_done[s])            # the slot's previous step no longer reads the buffers
            self.Q_dev[s][lo:hi].copy_(Q_host[lo:hi], non_blocking=True)
            self.C_dev[s][lo:hi].copy_(cand_host[lo:hi], non_blocking=True)
            if self.world > 1:                                      # slices → full replicated batch over NVLink
                dist.all_gather_into_tensor(self.Q_dev[s].view(self.n_queries * self.q_len, self.dim),
                                            self.Q_dev[s][lo:hi].reshape(self.slice * self.q_len, self.dim).clone(),
                                            group=self.in_group)
                dist.all_gather_into_tensor(self.C_dev[s], self.C_dev[s][lo:hi].clone(), group=self.in_group)
            self.ev_in[s].record(self.copy_stream)
        compute.wait_event(self.ev_in[s])
        pids, scores = self.ranker.rank_forward_batch(self.Q_dev[s], self.C_dev[s], depth=self.depth)
        self.out_pids[s].copy_(pids, non_blocking=True)
        self.out_scores[s].copy_(scores, non_blocking=True)
        self.ev_done[s].record(compute)
        return s

    def result(self, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Block until the step in ``slot`` has finished; → (pids [n_queries,k] int64, scores [n_queries,k] fp32) on the host."""
        self.ev_done[slot].synchronize()
        return self.out_pids[slot], self.out_scores[slot]
