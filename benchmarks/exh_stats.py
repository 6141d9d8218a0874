"""Per-role cycle accounting of maxsim_exhaustive_kernel (CBK_EXH_STATS=1 selects the instrumented instantiation):
    CBK_EXH_STATS=1 python benchmarks/exh_stats.py [nq ...]        (300 k documents, U[20,120] rows, fp16)"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from colbert_b200.ranking import ColbertRanker

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(7)
docs = 300000
doclens = torch.randint(20, 121, (docs,), generator=g, dtype=torch.int64)
total = int(doclens.sum())
store = torch.zeros(total + 512, 128, dtype=torch.float16, device=dev)
gg = torch.Generator(device=dev).manual_seed(8)
store[:total] = torch.nn.functional.normalize(torch.randn(total, 128, generator=gg, device=dev), dim=1).half()
ranker = ColbertRanker.from_store(store, doclens)
for nq in [int(x) for x in sys.argv[1:]] or (4, 8, 16):
    Q = torch.nn.functional.normalize(torch.randn(nq, 32, 128, generator=g), dim=2).to(dev)
    ranker.score_all(Q)
    ranker.score_all(Q)
    torch.cuda.synchronize()
