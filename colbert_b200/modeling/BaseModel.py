"""``BaseModel`` — the scoring operator and the multi-view switch, with the reference's names
(reference colbert/modeling/BaseModel.py:9-46).

Only the two members on the late-interaction scoring path are provided:

* ``BaseModel.score(Q, D, q_mask, d_mask)`` — all-pairs MaxSim, a ``@staticmethod`` exactly like the
  reference's so the class itself can be injected as ``model=`` into ``ColbertRanker``; it runs the
  fused sm_100a kernel (never materialising ``simmat``), not ``einsum``.
* ``get_representation(t, is_query)`` — the ``enable_multiview`` switch (keep the first
  ``q_view`` / ``d_view`` hidden states, project, L2-normalise).  The BERT encoder around it
  (``query`` / ``doc``, BaseModel.py:29-37) is out of scope.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import kernels
from .._lib import CBK_MAX_QLEN


class BaseModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.model = None
        self.linear = None
        self.args = None

    @property
    def encoder(self):
        return self.model

    def get_representation(self, t, is_query):
        """reference BaseModel.py:21-27.  With ``args.enable_multiview`` every query has exactly
        ``q_view`` rows and every document exactly ``d_view`` rows, all unit-norm, so "max over views"
        is the ordinary max over document rows in :meth:`score`."""
        if self.args.enable_multiview:
            mv = self.args.dense_multiview_args
            view_num = mv.q_view if is_query else mv.d_view
            t = t[:, :view_num, ...]
        t = self.linear(t)
        return torch.nn.functional.normalize(t, p=2, dim=2)

    @staticmethod
    def score(Q, D, q_mask, d_mask, *args, store_dtype: torch.dtype = torch.float16, **kwargs):
        """All-pairs MaxSim (reference BaseModel.py:39-46):

            scores[q, d] = Σ_m max_n (Q[q,m]·q_mask[q,m]) · (D[d,n]·d_mask[d,n])

        ``Q [q, m, h]``, ``D [d, n, h]``, ``q_mask [q, m]``, ``d_mask [d, n]`` → fp32 ``[q, d]``.
        Masks are multiplicative, so a masked slot contributes similarity 0 to the max — the same
        zero-floor the reference has.  Inputs must be CUDA tensors; the products are formed in fp32,
        rounded once to ``store_dtype`` (documents) and consumed by 16-bit tensor-core MMAs with fp32
        accumulation."""
        if not (Q.is_cuda and D.is_cuda):
            raise RuntimeError("colbert_b200 BaseModel.score runs on CUDA tensors only (no CPU path)")
        dev = D.device
        nq, m, h = Q.shape
        nd, n, h2 = D.shape
        assert h == h2, (Q.shape, D.shape)

        def as_rows(x):
            x = x.contiguous()
            return x if x.dtype in (torch.float16, torch.bfloat16, torch.float32) else x.float()

        def as_mask(x, rows):
            x = x.to(dev).contiguous().reshape(rows)
            return x if x.dtype in (torch.bool, torch.uint8, torch.int64, torch.float32) else x.float()

        Dm = kernels.mask_cast_rows(as_rows(D).reshape(nd * n, h), as_mask(d_mask, nd * n), store_dtype)
        Qm = kernels.mask_cast_rows(as_rows(Q.to(dev)).reshape(nq * m, h), as_mask(q_mask, nq * m), torch.float32)
        Qm = Qm.reshape(nq, m, h)
        doclens = torch.full((nd,), n, dtype=torch.int32, device=dev)
        pfxsum = torch.arange(0, (nd + 1) * n, n, dtype=torch.int64, device=dev)
        cand = torch.arange(nd, dtype=torch.int64, device=dev).repeat(nq)
        rowptr = torch.arange(0, (nq + 1) * nd, nd, dtype=torch.int64, device=dev)
        total = None
        for lo in range(0, m, CBK_MAX_QLEN):
            part = kernels.maxsim_rerank(Dm, pfxsum, doclens, [], Qm[:, lo: lo + CBK_MAX_QLEN].contiguous(), cand, rowptr)
            total = part if total is None else total.add_(part)
        return total.reshape(nq, nd)
