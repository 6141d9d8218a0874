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
    def __init__(self, ranker, n_queries: int, q_len: int, n_cand: int, depth: int = 10, slots: int = 2,
                 input_exchange: str = "p2p"):
        """``ranker``: ``ColbertRanker`` or ``ShardedColbertRanker``.  Every step carries ``n_queries`` queries
        of ``q_len`` rows and ``n_cand`` candidates each (equal-length lists).

        ``input_exchange`` (sharded stores): how the replicated inputs reach every GPU.  ``"allgather"``: each rank uploads
        its 1/world slice and the slices are all-gathered over NVLink (1/world of the PCIe traffic, but the NCCL kernel
        has to find SMs next to the persistent MaxSim kernel of the previous step).  ``"replicate"``: each rank uploads
        the whole batch through its own PCIe link — copy engines only, no SM and no collective on the input side.
        Measured on 8 x B200 (4096 queries x 1000 candidates per GPU, 16-bit queries / 32-bit pids, device step 15.5-16.3 ms):
        all-gather 17.4 ms per step end to end, replicate 21.0 ms (8 x 400 MB per step out of one host's memory) — hence
        the default.  ``"p2p"``: each rank uploads its slice into a SYMMETRIC buffer (torch.distributed._symmetric_memory) and
        pulls the other ranks' slices with peer-to-peer copies — copy engines over NVLink, no SM: unlike the NCCL kernel of
        "allgather" they run underneath the persistent MaxSim kernel of the previous step (8 x B200: 17.05 ms per step end to
        end against 17.4 ms, device step 16.2 ms; what is left is the widening of the 16-bit queries / 32-bit pids, two small
        kernels that do have to wait for SMs).  The default; falls back to "allgather" when symmetric memory cannot be
        set up."""
        self.ranker = ranker
        local = getattr(ranker, "local", ranker)
        self.device = local.device
        self.dim = local.dim
        sharded = hasattr(ranker, "local")                      # ShardedColbertRanker (ColbertRanker.rank is a method)
        self.world = ranker.world if sharded else 1
        self.rank = ranker.rank if sharded else 0
        self.group = ranker.group if sharded else None
        self.n_queries, self.q_len, self.n_cand = n_queries, q_len, n_cand
        assert input_exchange in ("allgather", "replicate", "p2p")
        self.input_exchange = input_exchange if self.world > 1 else "replicate"
        self.k = min(int(depth), n_cand)
        self.depth = depth
        self.slots = slots
        assert n_queries % self.world == 0, "the batch must split evenly over the ranks"
        self.slice = n_queries // self.world
        #: rows of the batch whose results THIS rank hands back (every rank holds the full replicated result on the
        #: device; each copies only its own 1/world slice to the host, together they return the whole batch)
        self.result_rows = (self.rank * self.slice, (self.rank + 1) * self.slice)
        # The tensor-core kernels round the query to fp16 on load, so fp16 queries on the wire give bit-identical
        # scores at half the bytes; the exceptions multiply differently and keep fp32 on the wire.
        self.fp16_wire_ok = bool(local.query_rounded_to_fp16)
        dev = self.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.Q_dev = [torch.empty((n_queries, q_len, self.dim), dtype=torch.float32, device=dev) for _ in range(slots)]
        self.C_dev = [torch.empty((n_queries, n_cand), dtype=torch.int64, device=dev) for _ in range(slots)]
        self.Q16_dev = [None] * slots           # staging for 16-bit queries / 32-bit pids on the wire, made on first use
        self.C32_dev = [None] * slots
        lo, hi = self.result_rows
        self.out_pids = [torch.empty((hi - lo, self.k), dtype=torch.int64).pin_memory() for _ in range(slots)]
        self.out_scores = [torch.empty((hi - lo, self.k), dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_done = [torch.cuda.Event() for _ in range(slots)]
        self._i = 0
        # the input all-gather runs on the copy stream, concurrently with the key all-gather of the previous step on
        # the compute stream: it needs its own communicator (collective call: every rank builds its pipeline)
        self._sym = None
        if self.world > 1 and self.input_exchange == "p2p":
            try:
                self._setup_p2p()
            except Exception as e:                              # no peer access / no symmetric-memory support on this box
                import warnings
                warnings.warn(f"RerankPipeline: symmetric-memory input exchange unavailable ({e!r}); using the NCCL all-gather")
                self.input_exchange = "allgather"
        self.in_group = (dist.new_group(ranks=list(range(self.world)))
                         if self.world > 1 and self.input_exchange == "allgather" else None)
        self.h2d_bytes_per_step = 0             # set by the first submit (depends on the dtypes handed in)
        self.d2h_bytes_per_step = n_queries * self.k * 12        # all ranks together: each its own slice
        self._wire = None

    def _setup_p2p(self) -> None:
        """One symmetric allocation per slot: [queries at up to 4 B per element | pids at up to 8 B]; typed views of it (mine
        and every peer's) are made on first use of a wire dtype."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        self._sym_q_bytes = self.n_queries * self.q_len * self.dim * 4
        self._sym_c_bytes = self.n_queries * self.n_cand * 8
        self._sym = symm_mem.empty((self.slots, self._sym_q_bytes + self._sym_c_bytes), dtype=torch.uint8, device=self.device)
        self._sym_hdl = symm_mem.rendezvous(self._sym, group)
        self._sym_peers = [self._sym_hdl.get_buffer(p, tuple(self._sym.shape), torch.uint8) for p in range(self.world)]
        self._sym_views = {}

    def _sym_view(self, s: int, which: str, dtype: torch.dtype, shape):
        """→ list over ranks of the typed view of slot ``s``'s query / pid region in that rank's symmetric buffer"""
        key = (s, which, dtype)
        if key not in self._sym_views:
            off = 0 if which == "q" else self._sym_q_bytes
            nbytes = int(torch.tensor([], dtype=dtype).element_size()) * int(torch.Size(shape).numel())
            self._sym_views[key] = [b[s, off: off + nbytes].view(dtype).view(shape) for b in self._sym_peers]
        return self._sym_views[key]

    def describe(self) -> str:
        up = {"allgather": f"every rank uploads 1/{self.world} of the batch (all-gathered over NVLink by NCCL)",
              "p2p": f"every rank uploads 1/{self.world} of the batch into symmetric memory and pulls the other slices with "
                     "peer-to-peer copies (copy engines over NVLink)",
              "replicate": "every rank uploads the whole batch through its own PCIe link"}[self.input_exchange]
        return ("colbert_b200.ranking.pipeline.RerankPipeline.submit/result (pinned host in, pinned host out, 2-slot stream "
                f"pipeline; wire dtypes {self._wire}; {up} and downloads 1/{self.world} of the result)")

    def _upload(self, s: int, host: torch.Tensor, full: torch.Tensor, lo: int, hi: int) -> None:
        """host[lo:hi] → full[lo:hi] (H2D), then — sharded — all-gather IN PLACE: the send buffer is this rank's own
        slice of the receive buffer, so no staging copy is made."""
        if self.input_exchange == "replicate":
            full.copy_(host, non_blocking=True)
            return
        full[lo:hi].copy_(host[lo:hi], non_blocking=True)
        flat = full.view(self.world, -1)
        dist.all_gather_into_tensor(flat, flat[self.rank], group=self.in_group)

    def submit(self, Q_host: torch.Tensor, cand_host: torch.Tensor, q_lens: Optional[torch.Tensor] = None,
               cand_rowptr: Optional[torch.Tensor] = None) -> int:
        """Queue one step; returns the slot to pass to :meth:`result`.  ``Q_host`` ``[n_queries, q_len, dim]`` fp32 or
        fp16 and ``cand_host`` ``[n_queries, n_cand]`` int64 or int32 (pids < 2^31), both pinned and identical on every
        rank.  16-bit queries / 32-bit pids halve the bytes on PCIe and NVLink and are widened on the device; the scores
        are bit-identical to the fp32 / int64 submission (the kernel rounds the query to fp16 itself).

        ``q_lens`` (host ``[n_queries]`` ints, optional): real row count of each query; the remaining rows of its
        ``q_len`` slots are padding (queries of different lengths in one batch — the reference's server strips the
        padding one query at a time, dense_server_client.py:44-46).
        ``cand_rowptr`` (host ``[n_queries + 1]`` int64, optional): RAGGED candidate lists — ``cand_host`` is then the
        flat concatenation (at most ``n_queries * n_cand`` pids, no list longer than ``n_cand``)."""
        assert Q_host.is_pinned() and cand_host.is_pinned(), "inputs must be pinned host tensors"
        assert Q_host.dtype in (torch.float32, torch.float16) and cand_host.dtype in (torch.int64, torch.int32)
        if cand_rowptr is not None:
            return self._submit_ragged(Q_host, cand_host, cand_rowptr, q_lens)
        if Q_host.dtype == torch.float16 and not self.fp16_wire_ok:
            raise ValueError("fp16 queries on the wire need a kernel that rounds the query to fp16 (dim multiple of 64, "
                             "no CBK_FLAG_BF16_NATIVE_MMA / CBK_FLAG_RERANK_GENERIC): pass fp32")
        s = self._i % self.slots
        self._i += 1
        compute = torch.cuda.current_stream(self.device)
        lo, hi = self.rank * self.slice, (self.rank + 1) * self.slice
        dev = self.device
        if self.input_exchange == "p2p":
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.ev_done[s])
                qv = self._sym_view(s, "q", Q_host.dtype, (self.n_queries, self.q_len, self.dim))
                cv = self._sym_view(s, "c", cand_host.dtype, (self.n_queries, self.n_cand))
                qv[self.rank][lo:hi].copy_(Q_host[lo:hi], non_blocking=True)          # my slice → my symmetric buffer
                cv[self.rank][lo:hi].copy_(cand_host[lo:hi], non_blocking=True)
                self._sym_hdl.barrier(channel=s)                                      # every rank's slice is in place
                for step in range(1, self.world):                                     # pull the others', staggered by rank
                    p = (self.rank + step) % self.world
                    plo, phi = p * self.slice, (p + 1) * self.slice
                    qv[self.rank][plo:phi].copy_(qv[p][plo:phi], non_blocking=True)
                    cv[self.rank][plo:phi].copy_(cv[p][plo:phi], non_blocking=True)
                # (no second barrier: a peer's pull of my slice precedes its scoring of this step, and my next write to this
                #  slot follows my scoring of this step, which ends in a collective that peer takes part in)
                self.Q_dev[s].copy_(qv[self.rank])                                    # widen (or copy) into the kernel's types
                self.C_dev[s].copy_(cv[self.rank])
                self.ev_in[s].record(self.copy_stream)
        else:
          with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_done[s])            # the slot's previous step no longer reads the buffers
            if Q_host.dtype == torch.float16:
                if self.Q16_dev[s] is None:
                    self.Q16_dev[s] = torch.empty((self.n_queries, self.q_len, self.dim), dtype=torch.float16, device=dev)
                self._upload(s, Q_host, self.Q16_dev[s], lo, hi)
                self.Q_dev[s].copy_(self.Q16_dev[s])                # exact widening
            else:
                self._upload(s, Q_host, self.Q_dev[s], lo, hi)
            if cand_host.dtype == torch.int32:
                if self.C32_dev[s] is None:
                    self.C32_dev[s] = torch.empty((self.n_queries, self.n_cand), dtype=torch.int32, device=dev)
                self._upload(s, cand_host, self.C32_dev[s], lo, hi)
                self.C_dev[s].copy_(self.C32_dev[s])
            else:
                self._upload(s, cand_host, self.C_dev[s], lo, hi)
            self.ev_in[s].record(self.copy_stream)
        if self._wire is None:
            self._wire = f"Q {str(Q_host.dtype).replace('torch.', '')}, pids {str(cand_host.dtype).replace('torch.', '')}"
            rows = self.n_queries if self.input_exchange == "replicate" else self.slice
            self.h2d_bytes_per_step = (rows * self.q_len * self.dim * Q_host.element_size()
                                       + rows * self.n_cand * cand_host.element_size()) * self.world
        compute.wait_event(self.ev_in[s])
        extra = {} if q_lens is None else {"q_lens": torch.as_tensor(q_lens, dtype=torch.int32)}
        pids, scores = self.ranker.rank_forward_batch(self.Q_dev[s], self.C_dev[s], depth=self.depth, **extra)
        r0, r1 = self.result_rows
        self.out_pids[s].copy_(pids[r0:r1], non_blocking=True)
        self.out_scores[s].copy_(scores[r0:r1], non_blocking=True)
        self.ev_done[s].record(compute)
        return s

    def _submit_ragged(self, Q_host, cand_host, cand_rowptr, q_lens) -> int:
        """Ragged lists: every rank uploads the whole (flat) batch itself — list boundaries do not fall on rank slices."""
        s = self._i % self.slots
        self._i += 1
        compute = torch.cuda.current_stream(self.device)
        rp = torch.as_tensor(cand_rowptr, dtype=torch.int64)
        total = int(rp[-1])
        assert rp.numel() == self.n_queries + 1 and total <= self.n_queries * self.n_cand
        assert int((rp[1:] - rp[:-1]).max()) <= self.n_cand, "a candidate list is longer than the pipeline was built for"
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_done[s])
            self.Q_dev[s].copy_(Q_host, non_blocking=True)                           # (widens fp16 on the way)
            flat = self.C_dev[s].view(-1)[:total]
            flat.copy_(cand_host.view(-1)[:total], non_blocking=True)
            rp_dev = rp.to(self.device, non_blocking=True)
            ql_dev = None if q_lens is None else torch.as_tensor(q_lens, dtype=torch.int32).to(self.device, non_blocking=True)
            self.ev_in[s].record(self.copy_stream)
        if self._wire is None:
            self._wire = f"Q {str(Q_host.dtype).replace('torch.', '')}, pids {str(cand_host.dtype).replace('torch.', '')}, ragged"
            self.h2d_bytes_per_step = (Q_host.numel() * Q_host.element_size() + total * cand_host.element_size()
                                       + rp.numel() * 8) * self.world
        compute.wait_event(self.ev_in[s])
        for t in (rp_dev, ql_dev):
            if t is not None:
                t.record_stream(compute)
        extra = {} if ql_dev is None else {"q_lens": ql_dev}
        pids, scores = self.ranker.rank_forward_batch(self.Q_dev[s], flat, rp_dev, depth=self.depth, max_cand=self.n_cand,
                                                      **extra)
        r0, r1 = self.result_rows
        self.out_pids[s].copy_(pids[r0:r1], non_blocking=True)
        self.out_scores[s].copy_(scores[r0:r1], non_blocking=True)
        self.ev_done[s].record(compute)
        return s

    def result(self, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Block until the step in ``slot`` has finished; → (pids [rows,k] int64, scores [rows,k] fp32) on the host for
        the queries ``result_rows`` of the batch (all of them on a single GPU; this rank's 1/world slice when sharded)."""
        self.ev_done[slot].synchronize()
        return self.out_pids[slot], self.out_scores[slot]


class GraphedRerank:
    """Low-batch latency path: host inputs → H2D → MaxSim → top-k → D2H captured ONCE in a CUDA graph and replayed per
    call (single GPU).  The reference's serving loop handles one query per call (``DenseRetrieverServer.retrieve`` →
    ``ColbertRetriever.search`` → ``rank_forward``, dense_server_client.py:36-49); here a call carries ``batch`` queries
    of up to ``q_len`` rows (``q_lens`` says how many are real) with up to ``n_cand`` candidates each, and costs one
    graph launch instead of four library calls + two copies.

        g = GraphedRerank(ranker, batch=8, q_len=32, n_cand=1000, depth=10)
        pids, scores = g(Q[8, 32, dim] fp32, cand[8, 1000] int64, q_lens=[...])     # host tensors in, host tensors out
    """

    def __init__(self, ranker, batch: int, q_len: int, n_cand: int, depth: int = 10):
        from .. import kernels
        self.ranker, self.batch, self.q_len, self.n_cand = ranker, batch, q_len, n_cand
        self.k = min(int(depth), n_cand)
        dev = self.device = ranker.device
        dim = ranker.dim
        self.h_Q = torch.zeros((batch, q_len, dim), dtype=torch.float32).pin_memory()
        self.h_C = torch.zeros((batch, n_cand), dtype=torch.int64).pin_memory()
        self.h_ql = torch.full((batch,), q_len, dtype=torch.int32).pin_memory()
        self.h_rp = torch.arange(0, (batch + 1) * n_cand, n_cand, dtype=torch.int64).pin_memory()
        self.h_pids = torch.empty((batch, self.k), dtype=torch.int64).pin_memory()
        self.h_scores = torch.empty((batch, self.k), dtype=torch.float32).pin_memory()
        self.d_Q, self.d_C = torch.zeros_like(self.h_Q, device=dev), torch.zeros_like(self.h_C, device=dev)
        self.d_ql, self.d_rp = torch.zeros_like(self.h_ql, device=dev), torch.zeros_like(self.h_rp, device=dev)
        self.stream = torch.cuda.Stream(device=dev)

        def body():
            self.d_Q.copy_(self.h_Q, non_blocking=True)
            self.d_C.copy_(self.h_C, non_blocking=True)
            self.d_ql.copy_(self.h_ql, non_blocking=True)
            self.d_rp.copy_(self.h_rp, non_blocking=True)
            flat = self.d_C.view(-1)
            scores = ranker.score_candidates(self.d_Q, flat, self.d_rp, self.d_ql)
            s, p = kernels.topk_per_query(scores, flat, self.d_rp, self.k, n_cand)
            self.h_pids.copy_(p, non_blocking=True)
            self.h_scores.copy_(s, non_blocking=True)

        with torch.cuda.stream(self.stream):
            for _ in range(2):                      # warm up outside capture: tensor maps, function attributes, allocator
                body()
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                body()

    def __call__(self, Q: torch.Tensor, cand: torch.Tensor, q_lens=None, cand_rowptr=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``Q`` ``[b ≤ batch, q_len, dim]`` fp32, ``cand`` ``[b, n ≤ n_cand]`` int64 (or flat + ``cand_rowptr`` for ragged
        lists) → (pids ``[b, k]``, scores ``[b, k]``) host tensors (views of the pinned result buffers: copy them if they
        must outlive the next call); lists shorter than k are padded with (-1, -inf)."""
        b = Q.size(0)
        assert b <= self.batch and Q.size(1) <= self.q_len
        self.h_Q[:b, : Q.size(1)].copy_(Q)
        if Q.size(1) < self.q_len:
            self.h_Q[:b, Q.size(1):].zero_()
        self.h_ql[:b] = Q.size(1) if q_lens is None else torch.as_tensor(q_lens, dtype=torch.int32)
        if cand_rowptr is None:
            n = cand.size(1)
            assert cand.size(0) == b and n <= self.n_cand
            self.h_C.view(-1)[: b * n].copy_(cand.reshape(-1))
            self.h_rp[: b + 1] = torch.arange(0, (b + 1) * n, n, dtype=torch.int64)
        else:
            rp = torch.as_tensor(cand_rowptr, dtype=torch.int64)
            assert rp.numel() == b + 1 and int(rp[-1]) <= self.batch * self.n_cand
            self.h_C.view(-1)[: int(rp[-1])].copy_(cand.reshape(-1)[: int(rp[-1])])
            self.h_rp[: b + 1] = rp
        self.h_rp[b + 1:] = self.h_rp[b]                       # unused queries: empty lists
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        self.stream.synchronize()
        return self.h_pids[:b], self.h_scores[:b]
