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


def _as_rows(x):
    x = x.contiguous()
    return x if x.dtype in (torch.float16, torch.bfloat16, torch.float32) else x.float()


def _as_mask(x, rows, dev):
    x = x.to(dev).contiguous().reshape(rows)
    return x if x.dtype in (torch.bool, torch.uint8, torch.int64, torch.float32) else x.float()


class MaxSimAllPairs(torch.autograd.Function):
    """``BaseModel.score`` with a backward pass, on the tcgen05 all-pairs kernels (``csrc/score_allpairs.cu``).

    forward: the multiplicative masks of BaseModel.py:41-42 are fused with the cast to 16 bits
    (``cbk_mask_cast_rows``), then ``cbk_score_allpairs_fwd`` forms Σ_m max_n without writing ``simmat``
    (BaseModel.py:43) and keeps the arg-max (the reference's ``indices``, l.44) for the backward.
    backward: ``cbk_score_allpairs_bwd`` — what autograd derives for l.41-45, with fixed-order sums.
    Masks get no gradient (the reference passes integer attention masks)."""

    @staticmethod
    def forward(ctx, Q, D, q_mask, d_mask, operand_dtype):
        dev = D.device
        nq, m, h = Q.shape
        nd, n, _ = D.shape
        qm = _as_mask(q_mask, nq * m, dev)
        dm = _as_mask(d_mask, nd * n, dev)
        Qp = kernels.mask_cast_rows(_as_rows(Q.detach().to(dev)).reshape(nq * m, h), qm, operand_dtype).reshape(nq, m, h)
        Dp = kernels.mask_cast_rows(_as_rows(D.detach()).reshape(nd * n, h), dm, operand_dtype).reshape(nd, n, h)
        need_grad = Q.requires_grad or D.requires_grad
        scores, argmax = kernels.score_allpairs_fwd(Qp, Dp, want_argmax=need_grad)
        if need_grad:
            ctx.save_for_backward(Qp, Dp, argmax, qm, dm)
            ctx.in_dtypes = (Q.dtype, D.dtype)
            ctx.q_device = Q.device
        return scores

    @staticmethod
    def backward(ctx, grad_scores):
        Qp, Dp, argmax, qm, dm = ctx.saved_tensors
        want_dq, want_dd = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dq, dd = kernels.score_allpairs_bwd(Qp, Dp, grad_scores.contiguous().float(), argmax, qm, dm, want_dq, want_dd)
        if dq is not None:
            dq = dq.to(device=ctx.q_device, dtype=ctx.in_dtypes[0])
        if dd is not None:
            dd = dd.to(ctx.in_dtypes[1])
        return dq, dd, None, None, None


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
        accumulation.

        Widths that are a multiple of 64 up to 1024 with at most 32 query rows — the training shape,
        colbert/modeling/colbert_model.py:87-95 — run the tcgen05 all-pairs kernel and are differentiable with
        respect to ``Q`` and ``D`` (:class:`MaxSimAllPairs`); other shapes are forward-only and go through the
        rerank kernel, one (query, document) pair per candidate."""
        if not (Q.is_cuda and D.is_cuda):
            raise RuntimeError("colbert_b200 BaseModel.score runs on CUDA tensors only (no CPU path)")
        dev = D.device
        nq, m, h = Q.shape
        nd, n, h2 = D.shape
        assert h == h2, (Q.shape, D.shape)
        if kernels.score_allpairs_supported(m, h):
            return MaxSimAllPairs.apply(Q, D, q_mask, d_mask, store_dtype)
        if torch.is_grad_enabled() and (Q.requires_grad or D.requires_grad):
            raise RuntimeError(f"BaseModel.score: no backward pass for m = {m}, h = {h} (needs m <= {CBK_MAX_QLEN} and h a "
                               "multiple of 64 up to 1024); detach the inputs to score without gradients")
        as_rows = _as_rows

        def as_mask(x, rows):
            return _as_mask(x, rows, dev)

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
