"""GPU-resident ``ColbertRanker`` — the host-side mirror of the reference's ranking API
(reference colbert/ranking/colbert_ranker.py:15-137, 238-241).

Same constructor, attributes and ``rank_forward`` contract as the reference class; the work the
reference does per query on the CPU + over PCIe (stride-bucket gather, cast, mask, einsum/max/sum,
un-permute, full sort) is one fused MaxSim launch and one top-k launch of libcolbert_b200.so over
a store that lives in HBM.  There is no torch-op or CPU fallback.

What changed relative to the reference and why:
  * ``self.tensor`` (the flat fp16 store, still ``num_embeddings + 512`` rows) lives on the GPU;
  * ``self.buffers`` is empty — the pinned / device staging buffers of ``_create_buffers``
    (≈1.9 GB at the reference's BSIZE) are not needed because nothing is staged;
  * ``self.views`` (the ``as_strided`` stride-views) are still offered, lazily, for API compatibility
    and for ``output_D_embedding``; scoring never reads through them — the kernel reads exactly
    ``doclen`` rows per document and reproduces the reference's zero-floor through a per-document
    flag (``doclen ∉ strides``, SURVEY.md §8 a12′);
  * ``rank_forward_batch`` scores many queries, each with its own candidate list, in one launch.
"""
from __future__ import annotations

import ctypes as C
from array import array
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import kernels
from ..indexing.index_manager import load_index_part
from ..indexing.loaders import get_parts, load_doclens

BSIZE = 1 << 14          # reference colbert_ranker.py:11 — most candidates one query may carry
TAIL_PAD_ROWS = 512      # reference colbert_ranker.py:62
DEVICE = "cuda"


def torch_percentile(tensor: torch.Tensor, p: int):
    """k-th smallest value with k = int(p·len/100), 1-indexed (reference colbert_ranker.py:238-241)."""
    assert p in range(1, 100 + 1)
    assert tensor.dim() == 1
    return tensor.kthvalue(int(p * tensor.size(0) / 100.0)).values.item()


def flatten(L):
    """reference colbert/utils/utils.py:133-134"""
    return [x for y in L for x in y]


def _is_builtin_scorer(model) -> bool:
    """True when ``model`` is None or this package's own ``BaseModel`` (class or instance): scoring then is the
    fused kernel.  Any other object with ``.score`` is an injected scorer and is called as the reference calls it."""
    if model is None:
        return True
    from ..modeling.BaseModel import BaseModel
    return model is BaseModel or isinstance(model, BaseModel) or (isinstance(model, type) and issubclass(model, BaseModel)
                                                                  and model.score is BaseModel.score)


class ColbertRanker:
    """Drop-in for the reference ``ColbertRanker`` (upstream ColBERT: ``IndexPart`` + ``IndexRanker``)."""

    def __init__(self, index_path: Optional[str], model=None, dim: Optional[int] = None,
                 device: Union[str, torch.device, None] = None, store_dtype: torch.dtype = torch.float16,
                 verbose: bool = False):
        self.device = torch.device(device if device is not None else DEVICE)
        if self.device.type != "cuda":
            raise RuntimeError("colbert_b200.ColbertRanker needs a CUDA device (sm_100a); there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.maxsim_dtype = torch.float32
        self.store_dtype = store_dtype
        # The reference's plugin seam (colbert_ranker.py:28,111): any object with .score(Q, D, q_mask, d_mask).
        # None or this package's BaseModel → the fused kernel; anything else is HONOURED: rank_forward then gathers
        # the stride-bucket tensors on the GPU (cbk_gather_rows) and calls model.score on them, bucket by bucket,
        # exactly as the reference does (see _scores_injected).  The batched / exhaustive entry points have no
        # counterpart in the reference and always run the fused kernel.
        self.model = model
        self._model_is_builtin = _is_builtin_scorer(model)
        if not self._model_is_builtin and not hasattr(model, "score"):
            raise TypeError("model= must offer .score(Q, D, q_mask, d_mask) (reference colbert_ranker.py:111)")
        self.verbose = verbose
        self.dim = dim
        # bf16 stores are multiplied as fp16 by default (see cbk_maxsim_rerank `flags`); set to
        # kernels._lib.CBK_FLAG_BF16_NATIVE_MMA to multiply bf16 directly
        self.kernel_flags = 0
        self.pid_base = 0           # first global pid of this store (non-zero only for a shard, see sharding.py)
        if index_path is not None:
            _, self.parts_paths, _ = get_parts(index_path)
            self.parts_doclens = load_doclens(index_path, flatten=False)
            self.doclens = flatten(self.parts_doclens)
            self.num_embeddings = sum(self.doclens)
            self.tensor = self._load_parts(dim)
            self.init_ranker()

    # -- alternative constructor for stores that are already in memory (synthetic benches, tests) ----
    @classmethod
    def from_tensors(cls, embeddings: torch.Tensor, doclens: Sequence[int], device=None, model=None,
                     store_dtype: Optional[torch.dtype] = None) -> "ColbertRanker":
        """``embeddings`` is ``[num_embeddings, dim]`` (host or device, fp16/bf16); the 512-row zero
        tail of the reference layout is appended here."""
        self = cls(None, model=model, dim=embeddings.size(1), device=device,
                   store_dtype=store_dtype or embeddings.dtype)
        self.parts_paths, self.parts_doclens = [], [list(map(int, doclens))]
        self.doclens = self.parts_doclens[0]
        self.num_embeddings = int(embeddings.size(0))
        assert sum(self.doclens) == self.num_embeddings
        store = torch.zeros(self.num_embeddings + TAIL_PAD_ROWS, self.dim, dtype=self.store_dtype, device=self.device)
        store[: self.num_embeddings].copy_(embeddings, non_blocking=True)
        self.tensor = store
        self.init_ranker()
        return self

    @classmethod
    def from_store(cls, store: torch.Tensor, doclens, model=None) -> "ColbertRanker":
        """Adopt (no copy) a device store that already has the reference layout
        ``[num_embeddings + 512, dim]``; ``doclens`` is a list or an int64 tensor."""
        assert store.is_cuda and store.dim() == 2 and store.is_contiguous()
        self = cls(None, model=model, dim=store.size(1), device=store.device, store_dtype=store.dtype)
        self.parts_paths, self.parts_doclens = [], []
        self.doclens = doclens
        self.num_embeddings = int(store.size(0)) - TAIL_PAD_ROWS
        self.tensor = store
        self.init_ranker()
        assert int(self.doclens_pfxsum[-1]) == self.num_embeddings, "doclens do not add up to the store size"
        return self

    @classmethod
    def from_flat(cls, flat_path: str, device=None, pid_lo: int = 0, pid_hi: Optional[int] = None,
                  model=None) -> "ColbertRanker":
        """Load (a pid range of) an index converted with ``indexing.flat_store.convert_index``: raw mmap → HBM.
        For a shard, pass its pid range and use the result as the ``local`` ranker of ``ShardedColbertRanker``
        (which sets ``pid_base`` and the corpus-wide strides)."""
        from ..indexing.flat_store import load_flat
        dev = torch.device(device if device is not None else DEVICE)
        if dev.type != "cuda":
            raise RuntimeError("colbert_b200.ColbertRanker needs a CUDA device (sm_100a); there is no CPU path")
        store, doclens, lo, _ = load_flat(flat_path, dev, pid_lo, pid_hi)
        self = cls.from_store(store, doclens, model=model)
        self.pid_base = int(lo)
        return self

    # -- reference colbert_ranker.py:61-73 ---------------------------------------------------------
    def _load_parts(self, dim, verbose=None):
        store = torch.zeros(self.num_embeddings + TAIL_PAD_ROWS, dim, dtype=self.store_dtype, device=self.device)
        row = 0
        for idx, filename in enumerate(self.parts_paths):
            n_rows = sum(self.parts_doclens[idx])
            part = load_index_part(filename, verbose=False)
            assert part.size(0) == n_rows and part.size(1) == dim, (filename, tuple(part.shape), n_rows, dim)
            store[row: row + n_rows].copy_(part.to(self.store_dtype))
            row += n_rows
        return store

    # -- reference colbert_ranker.py:31-43 ---------------------------------------------------------
    def init_ranker(self):
        self.doclens = torch.as_tensor(self.doclens, dtype=torch.int64)
        self.doclens_pfxsum = torch.zeros(self.doclens.numel() + 1, dtype=torch.int64)
        torch.cumsum(self.doclens, 0, out=self.doclens_pfxsum[1:])
        self.dim = self.tensor.size(-1)
        # every document has the same number of rows (enable_multiview: d_view embeddings per document)? → the kernel can
        # compute offsets as pid * d instead of looking them up (CBK_FLAG_FIXED_DOCLEN)
        dl_min, dl_max = int(self.doclens.min()), int(self.doclens.max())
        self._doclen_const = dl_min if dl_min == dl_max and dl_min > 0 else 0
        strides = [torch_percentile(self.doclens, p) for p in [25, 50, 75]]
        strides.append(self.doclens.max().item())
        self.strides = sorted(list(set(strides)))
        if self.verbose:
            print(f"#> Using strides {self.strides}..", flush=True)
        self.buffers = {}
        self._host_scratch = None       # (device bytes, pinned bytes) of the single-call path, grown on demand
        # device-side copies the kernels index by pid
        self._doclens_dev = self.doclens.to(torch.int32).to(self.device)
        self._pfxsum_dev = self.doclens_pfxsum.to(self.device)
        assert self.tensor.size(0) < 2 ** 31, "store exceeds 2^31-1 rows"

    @property
    def strides(self) -> List[int]:
        return self._strides

    @strides.setter
    def strides(self, value) -> None:
        """Everything derived from the stride list follows it: the ctypes copy the single-call path hands to the
        library, and the stride-views.  (``ShardedColbertRanker`` replaces a shard's strides with the corpus-wide
        list; a stale copy would apply the zero floor with the wrong list.)"""
        self._strides = [int(s) for s in value]
        self._strides_c = (C.c_int32 * max(1, len(self._strides)))(*self._strides)
        self._views = None

    @property
    def views(self) -> List[torch.Tensor]:
        """reference colbert_ranker.py:45-51 — zero-copy stride-views of the store, built on demand."""
        if self._views is None:
            self._views = self._create_views(self.tensor)
        return self._views

    def _create_views(self, tensor: torch.Tensor):
        out = []
        for stride in self.strides:
            rows = tensor.size(0) - stride + 1
            out.append(torch.as_strided(tensor, (rows, stride, self.dim), (self.dim, self.dim, 1)))
        return out

    @property
    def effective_flags(self) -> int:
        """``kernel_flags`` plus what the index itself implies: CBK_FLAG_FIXED_DOCLEN when every document has the same
        number of rows and that number is the only stride (so no floor can apply)."""
        flags = int(self.kernel_flags)
        if getattr(self, "_doclen_const", 0) and self.strides == [self._doclen_const]:
            flags |= kernels._lib.CBK_FLAG_FIXED_DOCLEN
        return flags

    @property
    def query_rounded_to_fp16(self) -> bool:
        """True when the kernel this ranker's calls resolve to rounds the query to fp16 on load — 16-bit queries on the wire
        (``RerankPipeline``) then give bit-identical scores.  Not the case for the generic kernel (fp32 arithmetic), the
        bf16-native flag, and bf16 stores on the tcgen05 kernels, which multiply the query as bf16 value + bf16 residual
        (16 significant bits: more than fp16 carries)."""
        flags = int(self.effective_flags)
        L = kernels._lib
        if flags & (L.CBK_FLAG_BF16_NATIVE_MMA | L.CBK_FLAG_RERANK_GENERIC) or self.dim % 64 != 0 or self.dim > 1024:
            return False
        bf16 = self.tensor.dtype == torch.bfloat16
        if self.dim == 128:
            return not (bf16 and flags & L.CBK_FLAG_RERANK_TCGEN05)
        return not (bf16 and 256 <= self.dim and not flags & L.CBK_FLAG_RERANK_KSPLIT)

    # -- the batched primitive everything else goes through -----------------------------------------
    def score_candidates(self, Q: torch.Tensor, cand_pids: torch.Tensor, cand_rowptr: torch.Tensor,
                         q_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 scores, one per candidate, in candidate order.  ``Q`` is ``[B, q_len, dim]`` fp32 on the
        device; query rows beyond 32 are scored in 32-row slices whose partial sums are added.  ``q_lens``
        (``[B]`` int32 on the device, optional): query b has only ``q_lens[b]`` real rows, the rest is padding."""
        B, q_len, dim = Q.shape
        flags = self.effective_flags
        if q_len <= kernels._lib.CBK_MAX_QLEN:
            return kernels.maxsim_rerank(self.tensor, self._pfxsum_dev, self._doclens_dev, self.strides, Q,
                                         cand_pids, cand_rowptr, flags=flags, pid_base=self.pid_base, q_lens=q_lens)
        total = None
        for lo in range(0, q_len, kernels._lib.CBK_MAX_QLEN):
            ql = None if q_lens is None else (q_lens - lo).clamp_(min=0)
            part = kernels.maxsim_rerank(self.tensor, self._pfxsum_dev, self._doclens_dev, self.strides,
                                         Q[:, lo: lo + kernels._lib.CBK_MAX_QLEN].contiguous(), cand_pids, cand_rowptr,
                                         flags=flags, pid_base=self.pid_base, q_lens=ql)
            total = part if total is None else total.add_(part)
        return total

    def rank_forward_batch(self, Q: torch.Tensor, cand_pids: torch.Tensor, cand_rowptr: Optional[torch.Tensor] = None,
                           depth: Optional[int] = 10, max_cand: Optional[int] = None,
                           q_lens: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Many queries, each with its own candidates, in one MaxSim launch + one top-k launch.

        ``Q``: ``[B, q_len, dim]`` fp32 (host — ideally pinned — or device).  ``cand_pids``: ``[B, n]``
        int64 for equal-length lists, or flat ``[N]`` with ``cand_rowptr`` ``[B+1]`` (ragged lists).
        ``q_lens``: optional ``[B]`` real row count of each query (queries of different lengths padded to ``q_len``
        rows — what the reference's server does one query at a time with ``keep_nonzero``,
        dense_server_client.py:44-46).
        → ``(pids [B, k] int64, scores [B, k] fp32)`` on the device, score-descending; k = depth
        (None: the longest list); shorter lists are padded with (-1, -inf)."""
        if q_lens is not None:
            q_lens = torch.as_tensor(q_lens).to(self.device, dtype=torch.int32, non_blocking=True).contiguous()
        Q = Q.to(self.device, dtype=self.maxsim_dtype, non_blocking=True).contiguous()
        B = Q.size(0)
        cand_pids = cand_pids.to(self.device, non_blocking=True)
        if cand_rowptr is None:
            assert cand_pids.dim() == 2 and cand_pids.size(0) == B
            n = cand_pids.size(1)
            cand_rowptr = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=self.device)
            max_cand = n
            cand_pids = cand_pids.reshape(-1)
        else:
            cand_rowptr = cand_rowptr.to(self.device, non_blocking=True)
            if max_cand is None:
                max_cand = int((cand_rowptr[1:] - cand_rowptr[:-1]).max().item())
        cand_pids = cand_pids.contiguous()
        scores = self.score_candidates(Q, cand_pids, cand_rowptr, q_lens)
        k = max_cand if depth is None else min(int(depth), max_cand)
        top_scores, top_pids = kernels.topk_per_query(scores, cand_pids, cand_rowptr, k, max_cand)
        return top_pids, top_scores

    def score_all(self, Q: torch.Tensor) -> torch.Tensor:
        """fp32 ``[B, n_docs]``: every document of this store against every query (``Q`` ``[B, q_len ≤ 32, dim]``
        fp32 on the device) — the query-batched tcgen05 kernel (dim 128), or, for a fixed-length
        fp16 store at another width (multi-view index, dim = 64 k up to 1024), the all-pairs kernel over the store seen as a
        padded batch.  The dim-128 kernel walks the store row by row, so documents without rows are taken out first and
        get the score the reference gives them (0, see tests/test_oracle_properties.py)."""
        if int(self.doclens.max()) == 0:
            return torch.zeros((Q.size(0), self.doclens.numel()), dtype=torch.float32, device=self.device)
        d = getattr(self, "_doclen_const", 0)
        if (self.dim != 128 and d and self.strides == [d] and self.tensor.dtype == torch.float16
                and kernels.score_allpairs_supported(Q.size(1), self.dim)):
            # A fixed-length (multi-view) store at a wide width — the author's 16 views x 768: the flat store IS the padded
            # batch [n_docs, d_view, dim] of BaseModel.score, so the tcgen05 all-pairs kernel scores it in place (no floor can
            # apply: the single stride is the document length).  The query is rounded to fp16 as everywhere else.
            n_docs = int(self.doclens.numel())
            Qp = kernels.mask_cast_rows(Q.reshape(-1, self.dim), None, torch.float16).reshape(Q.size(0), Q.size(1), self.dim)
            Dp = self.tensor[: n_docs * d].view(n_docs, d, self.dim)
            return kernels.score_allpairs_fwd(Qp, Dp, want_argmax=False)[0]
        if self.dim != 128:
            # Any other store the query-batched kernels do not take (ragged documents at a wide width, bf16 multi-view
            # stores): every document becomes a candidate of every query and the rerank kernel of that width scores them —
            # the store is read once per query instead of once per batch, which is the HBM-bound optimum for a single query.
            B, n_docs = Q.size(0), int(self.doclens.numel())
            cand = (torch.arange(n_docs, dtype=torch.int64, device=self.device) + self.pid_base).repeat(B)
            rowptr = torch.arange(0, (B + 1) * n_docs, n_docs, dtype=torch.int64, device=self.device)
            return self.score_candidates(Q, cand, rowptr).view(B, n_docs)
        if getattr(self, "_doc_end_bits", None) is None:
            nonempty = self.doclens > 0
            if bool(nonempty.all()):
                self._exh_cols, self._exh_pfxsum = None, self._pfxsum_dev
            else:
                idx = torch.nonzero(nonempty).flatten()
                self._exh_cols = idx.to(self.device)
                self._exh_pfxsum = torch.cat([self.doclens_pfxsum[idx], self.doclens_pfxsum[-1:]]).to(self.device)
            self._doc_end_bits = kernels.build_doc_end_bits(self._exh_pfxsum, self.tensor.size(0))
        dense = kernels.maxsim_exhaustive(self.tensor, self._exh_pfxsum, self._doc_end_bits, self.strides, Q,
                                          flags=self.kernel_flags & kernels._lib.CBK_FLAG_BF16_NATIVE_MMA)
        if self._exh_cols is None:
            return dense
        full = torch.zeros((Q.size(0), self.doclens.numel()), dtype=torch.float32, device=self.device)
        full[:, self._exh_cols] = dense
        return full

    def rank_exhaustive(self, Q: torch.Tensor, k: int = 1000) -> Tuple[torch.Tensor, torch.Tensor]:
        """No candidate generation: score the whole store and keep the top ``k`` per query.
        ``Q``: ``[B, q_len, dim]`` fp32 (host or device) → ``(pids [B,k] int64, scores [B,k] fp32)`` on the device."""
        Q = Q.to(self.device, dtype=self.maxsim_dtype, non_blocking=True).contiguous()
        k = min(int(k), int(self.doclens.numel()))
        scores, pids = kernels.topk_dense(self.score_all(Q), k, pid_base=self.pid_base)
        return pids, scores

    # -- reference colbert_ranker.py:75-137 ----------------------------------------------------------
    def rank_forward(self, Q, pids, views=None, depth=10, output_D_embedding=False):
        """``Q``: ``[1, dim, q_len]`` (as ``ColbertRetriever.search`` passes it, faiss_indexers.py:232-234);
        ``pids``: list or int64 tensor.  → ``(pids, scores)`` Python lists, score-descending, at most
        ``depth`` long; with ``output_D_embedding`` → ``(pids, D fp32 [depth, stride, dim], mask bool)``."""
        assert len(pids) > 0
        assert Q.size(0) in [1, len(pids)]
        if Q.size(0) != 1:
            # the reference accepts one query per candidate here but then only ever scores row 0 of
            # each bucket (colbert_ranker.py:111-112 `[0]`); that path has no defined meaning to mirror
            raise ValueError("rank_forward expects Q of shape [1, dim, q_len]")
        if len(pids) > BSIZE:
            raise ValueError(f"{len(pids)} candidates exceed BSIZE={BSIZE} (reference colbert_ranker.py:11)")
        if not isinstance(pids, torch.Tensor):
            pids = np.asarray(pids, dtype=np.int64)          # one conversion serves the range check and the call
        self._check_pids(pids)
        if not self._model_is_builtin:
            return self._rank_forward_injected(Q, pids, depth, output_D_embedding)
        if (not output_D_embedding and Q.device.type == "cpu" and Q.size(2) <= kernels._lib.CBK_MAX_QLEN
                and not (isinstance(pids, torch.Tensor) and pids.device.type != "cpu")):
            return self._rank_forward_host(Q, pids, depth)
        Qb = Q.to(self.device, dtype=self.maxsim_dtype).permute(0, 2, 1).contiguous()   # [1, q_len, dim]
        pids_t = torch.as_tensor(pids, dtype=torch.int64).to(self.device)
        n = pids_t.numel()
        rowptr = torch.tensor([0, n], dtype=torch.int64, device=self.device)
        scores = self.score_candidates(Qb, pids_t, rowptr)
        k = n if depth is None else min(int(depth), n)
        top_scores, top_pids = kernels.topk_per_query(scores, pids_t, rowptr, k, n)
        if not output_D_embedding:
            return top_pids[0].tolist(), top_scores[0].tolist()
        # output_D_embedding: the reference can only concatenate when every candidate fell into one
        # stride bucket (colbert_ranker.py:131-132), i.e. multi-view / fixed-length indexes
        local = (pids_t - self.pid_base).cpu()                       # doclens / pfxsum are indexed by LOCAL pid
        buckets = self._buckets(self.doclens[local])
        if int(buckets.min()) != int(buckets.max()):
            raise RuntimeError("output_D_embedding needs all candidates in one stride bucket "
                               f"(got buckets {sorted(set(buckets.tolist()))}); the reference fails here too")
        stride = self.strides[int(buckets[0])]
        D, mask = kernels.gather_rows(self.tensor, self._pfxsum_dev, self._doclens_dev,
                                      (top_pids[0] - self.pid_base).contiguous(), stride)
        return top_pids[0].tolist(), D, mask

    def _buckets(self, doclens: torch.Tensor) -> torch.Tensor:
        """reference colbert_ranker.py:90 — bucket g = #{s in strides : s < doclen}."""
        return (doclens.unsqueeze(1) > torch.tensor(self.strides, device=doclens.device).unsqueeze(0) + 1e-6).sum(-1)

    def _check_pids(self, pids) -> None:
        """The reference indexes ``doclens[pids]`` on the host and raises IndexError for a pid outside the index
        (colbert_ranker.py:88).  A shard ranker (CBK_FLAG_SKIP_FOREIGN_PIDS) legitimately sees foreign pids and skips
        the check; device-resident pid tensors are not pulled back for it (the kernel scores them NaN, which sorts
        last)."""
        if self.kernel_flags & kernels._lib.CBK_FLAG_SKIP_FOREIGN_PIDS:
            return
        if isinstance(pids, torch.Tensor):
            if pids.device.type != "cpu" or pids.numel() == 0:
                return
            lo, hi = int(pids.min()), int(pids.max())
        else:
            lo, hi = int(pids.min()), int(pids.max())
        n_docs = int(self._doclens_dev.numel())
        if lo < self.pid_base or hi >= self.pid_base + n_docs:
            raise IndexError(f"candidate pid out of range: [{lo}, {hi}] vs index pids "
                             f"[{self.pid_base}, {self.pid_base + n_docs})")

    def _rank_forward_injected(self, Q, pids, depth, output_D_embedding):
        """The reference's own flow for an INJECTED scorer (colbert_ranker.py:88-137): bucket the candidates by
        stride, gather each bucket's ``[n_g, stride_g, dim]`` fp32 tensor and length mask (``cbk_gather_rows``: the
        same rows the reference's stride-view ``index_select`` reads, on the GPU), call
        ``model.score(Q [1,q_len,dim], D, q_mask, d_mask)[0]``, un-permute, sort."""
        dev = self.device
        Qd = Q.to(dev, dtype=self.maxsim_dtype)
        pids_t = torch.as_tensor(pids, dtype=torch.int64).to(dev)
        local = pids_t - self.pid_base
        buckets = self._buckets(self._doclens_dev[local].to(torch.int64))
        scores = torch.empty(pids_t.numel(), dtype=torch.float32, device=dev)
        D_all, M_all, order = [], [], []
        q_mask = torch.ones((1, Qd.size(2)), dtype=torch.long, device=dev)
        for g, stride in enumerate(self.strides):
            sel = torch.nonzero(buckets == g).flatten()
            if sel.numel() == 0:
                continue
            D, mask = kernels.gather_rows(self.tensor, self._pfxsum_dev, self._doclens_dev, local[sel].contiguous(), stride)
            scores[sel] = self.model.score(Q=Qd.permute(0, 2, 1), D=D, q_mask=q_mask, d_mask=mask.to(torch.long))[0] \
                .to(device=dev, dtype=torch.float32)
            if output_D_embedding:
                D_all.append(D); M_all.append(mask); order.append(sel)
        n = pids_t.numel()
        k = n if depth is None else min(int(depth), n)
        rowptr = torch.tensor([0, n], dtype=torch.int64, device=dev)
        top_scores, top_pids = kernels.topk_per_query(scores, pids_t, rowptr, k, n)
        if not output_D_embedding:
            return top_pids[0].tolist(), top_scores[0].tolist()
        if len(D_all) != 1:
            raise RuntimeError("output_D_embedding needs all candidates in one stride bucket; the reference fails here too")
        pos = {int(p): i for i, p in enumerate(pids_t[order[0]].tolist())}
        idx = torch.tensor([pos[int(p)] for p in top_pids[0].tolist()], device=dev)
        return top_pids[0].tolist(), D_all[0][idx], M_all[0][idx]

    def _rank_forward_host(self, Q: torch.Tensor, pids, depth):
        """Host query + host pids → Python lists through ONE library call (cbk_rank_forward_host): a staged
        host→device copy, the MaxSim launch, the top-k launch and one device→host copy of the winners."""
        lib = kernels._lib.load()
        dim, q_len = Q.size(1), Q.size(2)
        q0 = Q[0]
        if q0.dtype != torch.float32:
            q0 = q0.float()
        if q0.t().is_contiguous():
            dim_major = 0                                   # a permuted view of a [q_len, dim] matrix
        else:
            q0, dim_major = q0.contiguous(), 1              # the reference's own layout
        if isinstance(pids, torch.Tensor):
            keep = pids.to(torch.int64).contiguous()
            pid_ptr, n = keep.data_ptr(), keep.numel()
        elif isinstance(pids, np.ndarray):
            keep = np.ascontiguousarray(pids, dtype=np.int64)
            pid_ptr, n = keep.ctypes.data, keep.size
        else:
            keep = array("q", pids)
            pid_ptr, n = keep.buffer_info()
        k = n if depth is None else min(int(depth), n)
        need = lib.cbk_rank_forward_scratch_bytes(n, q_len, dim, k)
        if self._host_scratch is None or self._host_scratch[0].numel() < need:
            cap = max(need, 1 << 20)
            self._host_scratch = (torch.zeros(cap, dtype=torch.uint8, device=self.device),
                                  torch.zeros(cap, dtype=torch.uint8).pin_memory())
        d_scr, h_scr = self._host_scratch
        out_pids = (C.c_int64 * k)()
        out_scores = (C.c_float * k)()
        store = self.tensor
        switch = torch.cuda.current_device() != self.device.index
        if switch:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(self.device)
        try:
            rc = lib.cbk_rank_forward_host(store.data_ptr(), kernels._lib.dtype_code(store.dtype), store.size(0), dim,
                                           self._pfxsum_dev.data_ptr(), self._doclens_dev.data_ptr(),
                                           self._doclens_dev.numel(), int(self.pid_base),
                                           C.cast(self._strides_c, C.c_void_p), len(self.strides), q0.data_ptr(), q_len,
                                           dim_major, pid_ptr, n, k, C.cast(out_pids, C.c_void_p),
                                           C.cast(out_scores, C.c_void_p), d_scr.data_ptr(), h_scr.data_ptr(),
                                           d_scr.numel(), int(self.effective_flags),
                                           C.c_void_p(kernels._lib.current_stream_ptr(self.device)))
        finally:
            if switch:
                torch.cuda.set_device(prev)
        kernels._lib.check("cbk_rank_forward_host", rc)
        return list(out_pids), list(out_scores)

    # upstream ColBERT name for the same call (IndexRanker.rank)
    def rank(self, Q, pids, views=None, depth=10):
        return self.rank_forward(Q, pids, views=views, depth=depth)


IndexRanker = ColbertRanker
