#!/bin/bash
for d in 0 1 2 4 8 3 7 15; do echo "== MFM_TC_DEBUG=$d"; MFM_TC_DEBUG=$d timeout 120 python scripts/gemm_bench.py 2>&1 | grep "path 1" | grep -E "att1_fc1|dcStar|dW11"; done
