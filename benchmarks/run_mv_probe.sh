#!/bin/bash
# multi-view rerank: looked-up path vs fixed-doclen path vs gather-only probe, d_view 8 and 16 (one line each)
for d in 8 16; do
  timeout 120 python benchmarks/rerank_micro.py --doclen $d --q-len $d --no-fixed
  timeout 120 python benchmarks/rerank_micro.py --doclen $d --q-len $d
  CBK_RERANK_PROBE=1 timeout 120 python benchmarks/rerank_micro.py --doclen $d --q-len $d
done
timeout 120 python benchmarks/rerank_micro.py --doclen 0 --q-len 32
CBK_RERANK_PROBE=1 timeout 120 python benchmarks/rerank_micro.py --doclen 0 --q-len 32
