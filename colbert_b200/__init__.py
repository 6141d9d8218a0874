"""colbert_b200 — B200-native (sm_100a) late-interaction scoring path of wuyaoxuehun/colbert.

Public surface (names follow the reference):
    colbert_b200.ranking.ColbertRanker          gather → MaxSim → top-k over an HBM-resident store
    colbert_b200.modeling.BaseModel.BaseModel    .score (all-pairs MaxSim) and the multi-view switch
    colbert_b200.indexing.loaders / index_manager   the {i}.pt + doclens.{i}.json layout
    colbert_b200.kernels                         tensor-level wrappers over the C ABI (include/colbert_b200.h)
"""
__version__ = "0.1.0"
