"""CPU PORT of the reference ranker — TEST / BASELINE INFRASTRUCTURE ONLY (see maxsim_oracle.py).

The reference's ranking path is Python that calls torch ops; the arithmetic itself lives in PyTorch
(third-party, pinned torch==1.10.0 in the reference's requirements.txt:252).  This module restates
that path with the SAME torch CPU ops in the SAME order, so that timing it on the GPU box's host
cores is a fair stand-in for "the reference CPU ranker" (the reference package itself cannot travel
to the GPU box).  It is validated against the numpy oracle and the golden vectors in
tests/test_oracle_golden.py::test_torch_port_matches_reference.

Op sequence per query (reference colbert/ranking/colbert_ranker.py:75-137, BaseModel.py:39-46):
  lookup doclens/offsets → bucket by stride → per bucket: index_select on the as_strided stride-view
  into a staging buffer → cast fp32 → arange mask → (mask-mul, einsum 'qmh,dnh->qdmn', max, sum)
  → concatenate, invert the permutation → sort descending → truncate to depth.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def maxsim_score(Q: torch.Tensor, D: torch.Tensor, q_mask: torch.Tensor, d_mask: torch.Tensor) -> torch.Tensor:
    """BaseModel.score (reference BaseModel.py:39-46), op for op."""
    D = D * d_mask[..., None]
    Q = Q * q_mask[..., None]
    sim = torch.einsum("qmh,dnh->qdmn", Q, D)
    return sim.max(-1).values.sum(-1)


class CpuRankerPort:
    """Reference ColbertRanker semantics over an in-memory CPU store (fp16 [T+512, dim])."""

    def __init__(self, store_fp16: torch.Tensor, doclens: Sequence[int], max_candidates: int = 1 << 14):
        self.store = store_fp16
        self.dim = store_fp16.size(1)
        self.doclens = torch.as_tensor(list(doclens) if not torch.is_tensor(doclens) else doclens, dtype=torch.int64)
        self.pfxsum = torch.zeros(self.doclens.numel() + 1, dtype=torch.int64)
        torch.cumsum(self.doclens, 0, out=self.pfxsum[1:])
        n = self.doclens.numel()
        kth = lambda p: self.doclens.kthvalue(int(p * n / 100.0)).values.item()          # l.238-241
        self.strides = sorted({kth(25), kth(50), kth(75), int(self.doclens.max())})       # l.36-40
        rows = store_fp16.size(0)
        self.views = [torch.as_strided(store_fp16, (rows - s + 1, s, self.dim), (self.dim, self.dim, 1))
                      for s in self.strides]                                              # l.45-51
        # staging buffers (l.53-59); sized for the batch actually used rather than BSIZE=16384
        self.buffers = [torch.zeros(max_candidates, s, self.dim, dtype=store_fp16.dtype) for s in self.strides]
        self._strides_t = torch.tensor(self.strides)

    def rank_forward(self, Q: torch.Tensor, pids, depth: Optional[int] = 10) -> Tuple[List[int], List[float]]:
        """Q: [1, dim, q_len] fp32; pids: list / int64 tensor."""
        assert len(pids) > 0
        Q = Q.contiguous().to(torch.float32)
        pids = torch.as_tensor(pids, dtype=torch.int64)
        doclens, offsets = self.doclens[pids], self.pfxsum[pids]                           # l.88
        bucket = (doclens.unsqueeze(1) > self._strides_t.unsqueeze(0) + 1e-6).sum(-1)      # l.90
        position = torch.arange(pids.numel())
        scores_parts, position_parts = [], []
        ones = torch.ones((1, Q.size(2)), dtype=torch.long)
        for g, stride in enumerate(self.strides):                                          # l.96
            sel = bucket == g
            n_g = int(sel.sum())
            if n_g == 0:
                continue
            D = torch.index_select(self.views[g], 0, offsets[sel], out=self.buffers[g][:n_g])   # l.105
            D = D.to(torch.float32)                                                        # l.107
            mask = (torch.arange(stride) + 1).unsqueeze(0) <= doclens[sel].unsqueeze(-1)   # l.108-109
            s = maxsim_score(Q.permute(0, 2, 1), D, ones, mask.to(torch.long))[0]          # l.111-112
            scores_parts.append(s)
            position_parts.append(position[sel])
        inverse = torch.cat(position_parts).sort().indices                                 # l.120
        scores = torch.cat(scores_parts)[inverse]                                          # l.122
        order = scores.sort(descending=True)                                               # l.128
        return pids[order.indices].tolist()[:depth], scores[order.indices].tolist()[:depth]
