from .colbert_ranker import ColbertRanker, torch_percentile  # noqa: F401
