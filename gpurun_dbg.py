import numpy as np, torch, sys
sys.path.insert(0, '.')
from colbert_b200 import _lib, synthetic
from colbert_b200.ranking import ColbertRanker
DEV = torch.device("cuda", 0)
for dt in (torch.float16, torch.bfloat16):
    index = synthetic.make_index(2025, 60_000, dim=128, lo=1, hi=180)
    emb = torch.from_numpy(index.emb).to(dt)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    B, n = 128, 1000
    Q = torch.from_numpy(synthetic.make_queries(1, B, 32, 128)).to(DEV)
    candh = synthetic.make_candidates(2, B, index.num_docs, n)
    cand = torch.from_numpy(candh).to(DEV).reshape(-1)
    rowptr = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=DEV)
    s_mma = ranker.score_candidates(Q, cand, rowptr)
    ranker.kernel_flags |= _lib.CBK_FLAG_RERANK_TCGEN05
    for rep in range(3):
        s_umma = ranker.score_candidates(Q, cand, rowptr)
        d = (s_umma - s_mma).abs()
        bad = torch.nonzero(d > 5e-3).flatten()
        print(dt, 'rep', rep, 'max diff', d.max().item(), 'n bad', bad.numel(), 'nan', torch.isnan(s_umma).sum().item())
        if bad.numel():
            b = bad[:12].cpu().numpy()
            dl = index.doclens[candh.reshape(-1)[b]]
            print('  idx', b.tolist()); print('  doclen', dl.tolist()); print('  umma', s_umma[bad[:12]].cpu().numpy().round(3).tolist()); print('  mma ', s_mma[bad[:12]].cpu().numpy().round(3).tolist())
            allbad = bad.cpu().numpy(); print('  bad doclen hist >128:', (index.doclens[candh.reshape(-1)[allbad]] > 128).mean(), ' idx%64 hist', np.bincount(allbad % 64, minlength=64).tolist())
