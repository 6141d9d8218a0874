"""The batch all-gather in front of the training-shaped score (reference colbert/training/training_utils.py:22-45).

``ColBERT_List_qa.forward`` (reference colbert/modeling/colbert_model.py:87-90) gathers ``Q``, ``D`` and their masks from
every rank before ``BaseModel.score``: each rank then scores the WHOLE batch (world x its own questions against world x its
own passages), and only its own slice of the gathered tensors carries a gradient.  This is plumbing around the operator —
``torch.distributed`` (NCCL over NVLink for CUDA tensors, gloo in the CPU tests) — kept with the reference's names so that
the training call reads the same; the scoring itself is ``colbert_b200.modeling.BaseModel.BaseModel.score``
(``cbk_score_allpairs_fwd`` / ``_bwd``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def distributed_concat(tensor: torch.Tensor, num_total_examples: Optional[int] = None, concat: bool = True):
    """reference training_utils.py:22-32 — all-gather ``tensor`` from every rank; the gathered copies carry no gradient.
    ``concat``: one tensor along dim 0 (truncated to ``num_total_examples`` when given), else the per-rank list."""
    world = dist.get_world_size()
    if tensor.is_contiguous() and tensor.dim() >= 1:
        n0 = tensor.size(0)
        flat = torch.empty((world * n0,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
        dist.all_gather_into_tensor(flat, tensor.detach())              # one collective into one buffer
        res: List[torch.Tensor] = list(flat.split(n0, dim=0))
    else:
        res = [torch.empty_like(tensor) for _ in range(world)]
        dist.all_gather(res, tensor.detach().contiguous())
    if concat:
        out = torch.cat(res, dim=0)
        return out[:num_total_examples] if num_total_examples else out
    return res


def collection_qd_masks(data: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """reference training_utils.py:35-45 — for each of ``[Q, q_mask, D, d_mask]``: the concatenation over ranks in which THIS
    rank's slot is the original tensor (so gradients flow back into the local encoder), every other slot a gathered copy."""
    rank = dist.get_rank()
    out = []
    for t in data:
        parts = distributed_concat(t, concat=False)
        parts[rank] = t
        out.append(torch.cat(parts, dim=0))
    return out
