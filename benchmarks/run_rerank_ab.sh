#!/bin/bash
# multi-view + headline kernel timings (one JSON line each)
timeout 120 python benchmarks/rerank_micro.py --doclen 8 --q-len 8 --dtype bf16
timeout 120 python benchmarks/rerank_micro.py --doclen 8 --q-len 8 --dtype fp16
timeout 120 python benchmarks/rerank_micro.py --doclen 8 --q-len 16 --dtype bf16
timeout 120 python benchmarks/rerank_micro.py --doclen 8 --q-len 32 --dtype bf16
timeout 120 python benchmarks/rerank_micro.py --doclen 0 --q-len 32 --dtype bf16
