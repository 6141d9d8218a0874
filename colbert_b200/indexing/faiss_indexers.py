"""``ColbertRetriever`` — the caller of the scoring path, with the reference's names
(reference colbert/indexing/faiss_indexers.py:161-235).

Only what reaches the hot path is here: ``load_index`` builds the HBM-resident ``ColbertRanker`` and the
``ColbertIndex`` post-processor, ``search`` keeps the reference's signature (one ``[q_len, dim]`` query →
``(pids, scores)`` lists), ``search_batch`` is the batched form (one MaxSim launch + one top-k launch for all
queries).  The faiss IVFPQ build/search of the reference is third-party and out of scope; the per-token ANN
search is injected as a callable (see ``ranking/colbert_index.py``).
"""
from __future__ import annotations

from typing import Optional

import torch

from ..ranking.colbert_index import ColbertIndex, Searcher
from ..ranking.colbert_ranker import ColbertRanker


class ColbertRetriever:
    def __init__(self, index_path=None, faiss_index_path=None, faiss_depth=None, nprobe=None, rank=4, partitions=None,
                 model=None, dim=None, m=64, nbits=8, searcher: Optional[Searcher] = None, device=None, **kwargs):
        self.index_path = index_path
        self.faiss_index_path = faiss_index_path
        self.faiss_index = None
        self.ranker = None
        self.dim = dim
        self.model = model
        self.faiss_depth = faiss_depth if faiss_depth is not None else 256     # faiss_indexers.py:173
        self.nprobe = nprobe if nprobe is not None else 64                     # faiss_indexers.py:174
        self.searcher = searcher
        self.device = device

    def load_index(self, ranker: Optional[ColbertRanker] = None):
        """reference faiss_indexers.py:178-183"""
        self.ranker = ranker if ranker is not None else ColbertRanker(self.index_path, model=self.model, dim=self.dim,
                                                                      device=self.device)
        if self.searcher is None:
            raise RuntimeError("ColbertRetriever needs a `searcher` callable for candidate generation "
                               "(the reference uses a faiss IVFPQ index, which is outside this library)")
        self.faiss_index = ColbertIndex(self.ranker, self.searcher, nprobe=self.nprobe)

    def search(self, query: torch.Tensor, topk_doc=None, faiss_depth=None, **kwargs):
        """reference faiss_indexers.py:224-235 — ``query`` is ``[q_len, dim]``; → ``(pids, scores)`` lists."""
        Q = query
        assert len(Q.shape) == 2
        faiss_depth = self.faiss_depth if faiss_depth is None else faiss_depth
        pids = self.faiss_index.retrieve(faiss_depth, Q=query.unsqueeze(0))[0]
        Q = Q.unsqueeze(0).permute(0, 2, 1)                      # [1, dim, q_len], as the reference passes it
        return self.ranker.rank_forward(Q, pids, depth=topk_doc)

    def search_batch(self, queries: torch.Tensor, topk_doc: int = 10, faiss_depth=None):
        """``queries`` ``[B, q_len, dim]`` → ``(pids [B,k], scores [B,k])`` on the device; the candidate lists never
        leave the GPU between candidate generation and scoring."""
        faiss_depth = self.faiss_depth if faiss_depth is None else faiss_depth
        pids, rowptr = self.faiss_index.retrieve_csr(faiss_depth, queries)
        return self.ranker.rank_forward_batch(queries, pids, rowptr, depth=topk_doc,
                                              max_cand=min(queries.size(1) * faiss_depth, 1 << 14))
