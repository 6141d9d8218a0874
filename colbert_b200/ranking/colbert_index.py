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
    def __init__(self, ranker, searcher: Searcher, nprobe=None):
        self.ranker = ranker
        self.searcher = searcher
        self.nprobe = nprobe
        self.emb2pid = None
        self.build_emb2pid()

    def build_emb2pid(self):
        """reference colbert_ranker.py:163-174 — int32 ``[num_embeddings]``, row → pid (built on the device)."""
        self.emb2pid = kernels.build_emb2pid(self.ranker._pfxsum_dev)

    def queries_to_embedding_ids(self, faiss_depth: int, Q: torch.Tensor) -> torch.Tensor:
        """reference colbert_ranker.py:183-210 — flatten ``Q [B, q_len, dim]`` to rows, search each row."""
        B, q_len, dim = Q.shape
        ids = self.searcher(Q.reshape(B * q_len, dim).to(self.ranker.device, dtype=torch.float32), int(faiss_depth))
        return ids.to(torch.int64).reshape(B, q_len * faiss_depth).contiguous()

    def embedding_ids_to_pids(self, embedding_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """reference colbert_ranker.py:212-229 — → CSR ``(pids, rowptr)`` on the device: per query the sorted
        unique pids (the reference returns the same set as a Python list in ``set`` order)."""
        return kernels.embedding_ids_to_pids(embedding_ids.to(self.ranker.device).contiguous(), self.emb2pid)

    def retrieve_csr(self, faiss_depth: int, Q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.embedding_ids_to_pids(self.queries_to_embedding_ids(faiss_depth, Q))

    def retrieve(self, faiss_depth: int, Q: torch.Tensor, verbose: bool = False) -> List[List[int]]:
        """reference colbert_ranker.py:176-181 — list (per query) of candidate pid lists."""
        pids, rowptr = self.retrieve_csr(faiss_depth, Q)
        rp = rowptr.tolist()
        flat = pids[: rp[-1]].tolist()
        return [flat[rp[b]: rp[b + 1]] for b in range(len(rp) - 1)]
