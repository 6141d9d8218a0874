"""``ColbertIndex`` — candidate generation around an injected ANN searcher, with the reference's names
(reference colbert/ranking/colbert_ranker.py:140-235).

The reference searches a faiss IVFPQ index per query token (third-party, out of scope — SURVEY.md §2 row 5)
and then, on the CPU, maps embedding ids to pids through ``emb2pid`` and removes duplicates per query with
Python ``set`` (a ``Pool(16)`` when the batch is large).  Here the search itself is whatever callable the user
injects (faiss, cuVS, a brute-force scan …) and the post-processing — the part that feeds the scoring
path — runs on the GPU and emits the CSR candidate lists ``cbk_maxsim_rerank`` consumes directly.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch

from .. import kernels

# searcher(Q_flat fp32 [n_rows, dim] (device), depth) -> embedding ids int64 [n_rows, depth] (device; -1 = none)
Searcher = Callable[[torch.Tensor, int], torch.Tensor]


class ColbertIndex:
    """Two ways to build it:

    * ``ColbertIndex(ranker, searcher, nprobe=None)`` — around a ranker that is already in HBM;
    * ``ColbertIndex(index_path, faiss_index_path, nprobe, rank=None)`` — the reference's own signature
      (colbert_ranker.py:141): ``emb2pid`` is built from the ``doclens.*.json`` of ``index_path`` like the
      reference's ``build_emb2pid`` does, on device ``rank`` (default: the current device).  ``faiss_index_path`` is a
      searcher callable, or the path of a faiss index that is opened with the ``faiss`` package when it is installed
      (third-party, not part of this library; without it the constructor raises a clear error).
    """

    def __init__(self, ranker_or_index_path, searcher_or_faiss_index_path=None, nprobe=None, rank=None):
        self.nprobe = nprobe
        self.emb2pid = None
        if isinstance(ranker_or_index_path, (str, bytes)) or hasattr(ranker_or_index_path, "__fspath__"):
            import os
            self.index_path = os.fspath(ranker_or_index_path)
            self.faiss_index_path = searcher_or_faiss_index_path
            self.ranker = None
            self.device = torch.device("cuda", torch.cuda.current_device() if rank is None else int(rank))
            self.pid_base = 0
            self.searcher = self._open_searcher(searcher_or_faiss_index_path, nprobe)
        else:
            self.ranker = ranker_or_index_path
            self.index_path = self.faiss_index_path = None
            local = getattr(self.ranker, "local", self.ranker)      # a ShardedColbertRanker wraps the shard's ranker
            self.device = local.device
            self.pid_base = int(getattr(local, "pid_base", 0))
            self.searcher = searcher_or_faiss_index_path
        self.build_emb2pid()

    @staticmethod
    def _open_searcher(faiss_index_path, nprobe) -> Searcher:
        if callable(faiss_index_path):
            return faiss_index_path
        try:
            import faiss                                      # third-party ANN library, optional
        except ImportError as exc:
            raise RuntimeError("ColbertIndex(index_path, faiss_index_path, ...) needs the `faiss` package to open "
                               f"{faiss_index_path!r}; pass a searcher callable instead (candidate generation is "
                               "outside this library)") from exc
        index = faiss.read_index(faiss_index_path)
        if nprobe is not None:
            index.nprobe = nprobe

        def search(Q_rows: torch.Tensor, depth: int) -> torch.Tensor:
            _, ids = index.search(Q_rows.float().cpu().numpy(), depth)
            return torch.from_numpy(ids).to(Q_rows.device)
        return search

    def build_emb2pid(self):
        """reference colbert_ranker.py:163-174 — int32 ``[num_embeddings]``, row → pid (built on the device).
        Rows of a shard map to LOCAL pids here; ``embedding_ids_to_pids`` adds the shard's ``pid_base`` so that what
        it emits are the GLOBAL pids the scoring calls expect."""
        if self.ranker is not None:
            pfx = getattr(self.ranker, "local", self.ranker)._pfxsum_dev
        else:
            from ..indexing.loaders import load_doclens
            doclens = torch.as_tensor(load_doclens(self.index_path, flatten=True), dtype=torch.int64)
            pfx = torch.zeros(doclens.numel() + 1, dtype=torch.int64)
            torch.cumsum(doclens, 0, out=pfx[1:])
            pfx = pfx.to(self.device)
        self.emb2pid = kernels.build_emb2pid(pfx)

    def queries_to_embedding_ids(self, faiss_depth: int, Q: torch.Tensor) -> torch.Tensor:
        """reference colbert_ranker.py:183-210 — flatten ``Q [B, q_len, dim]`` to rows, search each row."""
        B, q_len, dim = Q.shape
        ids = self.searcher(Q.reshape(B * q_len, dim).to(self.device, dtype=torch.float32), int(faiss_depth))
        return ids.to(torch.int64).reshape(B, q_len * faiss_depth).contiguous()

    def embedding_ids_to_pids(self, embedding_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """reference colbert_ranker.py:212-229 — → CSR ``(pids, rowptr)`` on the device: per query the sorted
        unique pids (the reference returns the same set as a Python list in ``set`` order)."""
        pids, rowptr = kernels.embedding_ids_to_pids(embedding_ids.to(self.device).contiguous(), self.emb2pid)
        if self.pid_base:
            pids += self.pid_base          # (entries past rowptr[-1] are unused scratch)
        return pids, rowptr

    def retrieve_csr(self, faiss_depth: int, Q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.embedding_ids_to_pids(self.queries_to_embedding_ids(faiss_depth, Q))

    def retrieve(self, faiss_depth: int, Q: torch.Tensor, verbose: bool = False) -> List[List[int]]:
        """reference colbert_ranker.py:176-181 — list (per query) of candidate pid lists."""
        pids, rowptr = self.retrieve_csr(faiss_depth, Q)
        rp = rowptr.tolist()
        flat = pids[: rp[-1]].tolist()
        return [flat[rp[b]: rp[b + 1]] for b in range(len(rp) - 1)]
