"""The batch all-gather in front of the training-shaped score (reference colbert/training/training_utils.py:22-45) over gloo,
world size 2: layout of the gathered batch, and gradients reaching only the local slice."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from colbert_b200.training.training_utils import collection_qd_masks, distributed_concat
        torch.manual_seed(100 + rank)
        Q = torch.randn(3, 4, 8, requires_grad=True)
        D = torch.randn(6, 5, 8, requires_grad=True)
        qm = torch.ones(3, 4, dtype=torch.long)
        dm = (torch.arange(5)[None, :] < torch.tensor([[5], [4], [3], [2], [1], [5]])).long()
        gQ, gqm, gD, gdm = collection_qd_masks([Q, qm, D, dm])
        assert gQ.shape == (3 * world, 4, 8) and gD.shape == (6 * world, 5, 8) and gqm.shape == (3 * world, 4)
        assert torch.equal(gQ[3 * rank: 3 * rank + 3], Q) and torch.equal(gD[6 * rank: 6 * rank + 6], D)
        other = 1 - rank
        torch.manual_seed(100 + other)
        assert torch.equal(gQ[3 * other: 3 * other + 3], torch.randn(3, 4, 8))
        # the reference's score on the gathered batch (BaseModel.py:41-45), differentiated: only the local slot has a gradient
        sim = torch.einsum("qmh,dnh->qdmn", gQ * gqm[..., None], gD * gdm[..., None])
        sim.max(-1)[0].sum(-1).sum().backward()
        assert Q.grad is not None and D.grad is not None and float(Q.grad.abs().sum()) > 0
        cat = distributed_concat(torch.full((2,), float(rank)), num_total_examples=3)
        assert cat.tolist() == [0.0, 0.0, 1.0]
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_collection_qd_masks_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29571, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
