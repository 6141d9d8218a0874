"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups.  The two device stages of
ShardedColbertRanker (local scoring + top-k keys, merge) are replaced by the numpy oracle; shard
planning, global-pid arithmetic, the packed-key exchange through all_gather_into_tensor and the
"same answer on every rank, equal to the single-process answer" contract are what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colbert_b200 import synthetic
from colbert_b200.sharding import ShardedColbertRanker, owner_of, plan_shards
from oracle import maxsim_oracle as O


def test_plan_shards_balances_tokens():
    rng = np.random.default_rng(0)
    dl = torch.from_numpy(rng.integers(1, 181, size=10_000))
    pf = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(dl, 0)])
    for world in (1, 2, 3, 4, 8):
        b = plan_shards(pf, world)
        assert b[0] == 0 and b[-1] == 10_000 and len(b) == world + 1 and b == sorted(b)
        tokens = [int(pf[b[r + 1]] - pf[b[r]]) for r in range(world)]
        assert max(tokens) - min(tokens) <= 2 * 180
    # degenerate: more ranks than documents
    b = plan_shards(torch.tensor([0, 5, 9]), 4)
    assert b[0] == 0 and b[-1] == 2 and b == sorted(b)
    own = owner_of(torch.tensor([0, 4999, 9999]), plan_shards(pf, 2))
    assert own.tolist()[0] == 0 and own.tolist()[-1] == 1


class OracleShardedRanker(ShardedColbertRanker):
    """CPU stand-in: the oracle scores this rank's pid range; keys are packed exactly as the device does."""

    def __init__(self, index, bounds, strides, group=None):
        super().__init__(None, bounds[dist.get_rank()], strides, group)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        pf = O.doclens_pfxsum(index.doclens)
        self.lo, self.hi = lo, hi
        self.store = O.pad_store(index.emb[pf[lo]: pf[hi]])
        self.dl = index.doclens[lo:hi]
        self.pf = O.doclens_pfxsum(self.dl)

    def _local_topk_keys(self, Q, cand_pids, cand_rowptr, k, max_cand):
        Qn, pids, rp = Q.numpy(), cand_pids.numpy(), cand_rowptr.numpy()
        out = np.zeros((Qn.shape[0], k), dtype=np.uint64)
        for b in range(Qn.shape[0]):
            p = pids[rp[b]: rp[b + 1]]
            mine = (p >= self.lo) & (p < self.hi)
            if mine.any():
                s = O.maxsim_exact(self.store, self.dl, self.pf, self.strides, Qn[b], p[mine] - self.lo)
                keys = np.sort(O.pack_keys(s, p[mine]))[::-1][:k]
                out[b, : keys.shape[0]] = keys
        return torch.from_numpy(out.view(np.int64))

    def _local_num_docs(self):
        return self.hi - self.lo

    def _local_exhaustive_keys(self, Q, k):
        # rank_exhaustive asks for k_local = min(k, documents of this shard) keys and pads the list itself
        assert k <= self.hi - self.lo
        Qn = Q.numpy()
        out = np.zeros((Qn.shape[0], k), dtype=np.uint64)
        pids = np.arange(self.lo, self.hi, dtype=np.int64)
        for b in range(Qn.shape[0]):
            s = O.maxsim_exact(self.store, self.dl, self.pf, self.strides, Qn[b], pids - self.lo)
            keys = np.sort(O.pack_keys(s, pids))[::-1][:k]
            out[b, : keys.shape[0]] = keys
        return torch.from_numpy(out.view(np.int64))

    def _merge(self, gathered, k):
        g = gathered.numpy().view(np.uint64)                       # [W, B, k]
        W, B, kin = g.shape
        flat = np.transpose(g, (1, 0, 2)).reshape(B, W * kin)
        top = np.sort(flat, axis=1)[:, ::-1][:, :k]
        scores, pids = O.unpack_keys(top)
        return torch.from_numpy(pids.copy()), torch.from_numpy(scores.copy())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        index = synthetic.make_index(77, 600, dim=128, lo=1, hi=60)
        pf = torch.from_numpy(O.doclens_pfxsum(index.doclens))
        bounds = plan_shards(pf, world)
        strides = O.compute_strides(index.doclens)
        ranker = OracleShardedRanker(index, bounds, strides)
        B, n, k = 5, 90, 7
        Q = synthetic.make_queries(78, B, 32, 128)
        cand = synthetic.make_candidates(79, B, index.num_docs, n)
        cand[4, :] = np.arange(n)                                   # one query whose candidates all sit on rank 0
        pids, scores = ranker.rank_forward_batch(torch.from_numpy(Q), torch.from_numpy(cand), depth=k)
        ret[rank] = (pids.numpy().copy(), scores.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rank_equals_single_process(world):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    index = synthetic.make_index(77, 600, dim=128, lo=1, hi=60)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    strides = O.compute_strides(index.doclens)
    B, n, k = 5, 90, 7
    Q = synthetic.make_queries(78, B, 32, 128)
    cand = synthetic.make_candidates(79, B, index.num_docs, n)
    cand[4, :] = np.arange(n)
    for r in range(world):
        pids, scores = ret[r]
        assert np.array_equal(pids, ret[0][0]) and np.array_equal(scores, ret[0][1])      # identical on every rank
    for b in range(B):
        ref = O.maxsim_exact(store, index.doclens, pf, strides, Q[b], cand[b])
        rp, rs = O.topk_desc(ref, cand[b], k)
        assert ret[0][0][b].tolist() == rp.tolist()
        np.testing.assert_allclose(ret[0][1][b], rs, rtol=1e-6)


def _worker_exhaustive(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        index = synthetic.make_index(81, 120, dim=128, lo=1, hi=40)
        pf = torch.from_numpy(O.doclens_pfxsum(index.doclens))
        ranker = OracleShardedRanker(index, plan_shards(pf, world), O.compute_strides(index.doclens))
        Q = synthetic.make_queries(82, 3, 32, 128)
        pids, scores = ranker.rank_exhaustive(torch.from_numpy(Q), k=50)      # k larger than a shard of ~40 documents
        ret[rank] = (pids.numpy().copy(), scores.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_sharded_exhaustive_equals_single_process():
    """rank_exhaustive over 3 shards of ~40 documents with k = 50: every shard pads its list, the merged top-k equals
    the single-process ranking of the whole corpus on every rank."""
    world = 3
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_exhaustive, args=(world, port, ret), nprocs=world, join=True)
    index = synthetic.make_index(81, 120, dim=128, lo=1, hi=40)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    strides = O.compute_strides(index.doclens)
    Q = synthetic.make_queries(82, 3, 32, 128)
    allp = np.arange(index.num_docs, dtype=np.int64)
    for r in range(world):
        assert np.array_equal(ret[r][0], ret[0][0]) and np.array_equal(ret[r][1], ret[0][1])
    for b in range(3):
        ref = O.maxsim_exact(store, index.doclens, pf, strides, Q[b], allp)
        rp, rs = O.topk_desc(ref, allp, 50)
        assert ret[0][0][b].tolist() == rp.tolist()
        np.testing.assert_allclose(ret[0][1][b], rs, rtol=1e-6)
