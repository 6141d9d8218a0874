"""Upstream-ColBERT spelling of the scoring operator: ``colbert_score(Q, D_padded, D_mask)``.

This fork folded the operator into ``BaseModel.score`` (reference colbert/modeling/BaseModel.py:39-46); code written
against upstream ColBERT calls ``ModelInference.colbert_score`` / ``colbert_score`` instead: one query against many
padded documents (``Q.size(0) == 1``) or query *i* against document *i* (``Q.size(0) == D_padded.size(0)``).  Both shapes
run the same fused MaxSim kernel as the ranker; nothing is computed with torch ops.
"""
from __future__ import annotations

import torch

from .. import kernels
from .._lib import CBK_MAX_QLEN
from .BaseModel import BaseModel


def colbert_score(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor, store_dtype: torch.dtype = torch.float16
                  ) -> torch.Tensor:
    """``Q [1 | B, q_len, dim]``, ``D_padded [B, doclen, dim]``, ``D_mask [B, doclen]`` (or ``[B, doclen, 1]``) → fp32 ``[B]``:
    Σ over query rows of the max over the unmasked document rows.  Masked rows count as similarity 0, as in
    ``BaseModel.score``."""
    assert Q.dim() == 3 and D_padded.dim() == 3 and Q.size(0) in (1, D_padded.size(0)), (Q.shape, D_padded.shape)
    if not D_padded.is_cuda:
        raise RuntimeError("colbert_b200 colbert_score runs on CUDA tensors only (no CPU path)")
    B, n, h = D_padded.shape
    D_mask = D_mask.reshape(B, n)
    q_mask = torch.ones(Q.shape[:2], dtype=torch.float32, device=D_padded.device)
    if Q.size(0) == 1:
        return BaseModel.score(Q, D_padded, q_mask, D_mask, store_dtype=store_dtype)[0]
    # pairwise: every query owns exactly one candidate — the same launch shape as a rerank with one-element lists
    dev = D_padded.device
    mask = D_mask.to(dev).contiguous().reshape(B * n)
    if mask.dtype not in (torch.bool, torch.uint8, torch.int64, torch.float32):
        mask = mask.float()
    rows = D_padded.contiguous().reshape(B * n, h)
    if rows.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        rows = rows.float()
    Dm = kernels.mask_cast_rows(rows, mask, store_dtype)
    Qf = Q.to(dev, dtype=torch.float32).contiguous()
    doclens = torch.full((B,), n, dtype=torch.int32, device=dev)
    pfxsum = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=dev)
    cand = torch.arange(B, dtype=torch.int64, device=dev)
    rowptr = torch.arange(0, B + 1, dtype=torch.int64, device=dev)
    total = None
    for lo in range(0, Q.size(1), CBK_MAX_QLEN):
        part = kernels.maxsim_rerank(Dm, pfxsum, doclens, [], Qf[:, lo: lo + CBK_MAX_QLEN].contiguous(), cand, rowptr)
        total = part if total is None else total.add_(part)
    return total


class ModelInference:
    """Name upstream code imports; only the scoring member is on this path (the encoder is out of scope)."""

    def __init__(self, colbert=None, amp: bool = False):
        self.colbert, self.amp = colbert, amp

    @staticmethod
    def colbert_score(Q, D_padded, D_mask, **kwargs):
        return colbert_score(Q, D_padded, D_mask, **kwargs)

    def score(self, Q, D, q_mask=None, d_mask=None):
        q_mask = torch.ones(Q.shape[:2], device=D.device) if q_mask is None else q_mask
        d_mask = torch.ones(D.shape[:2], device=D.device) if d_mask is None else d_mask
        return BaseModel.score(Q, D, q_mask, d_mask)
