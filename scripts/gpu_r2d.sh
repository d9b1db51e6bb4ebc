#!/bin/bash
timeout 120 python scripts/lstm_trace.py 2>&1 | tee gpurun_out/lstm_trace.txt | grep -v "^  warp  [1235679]\|^  warp 1[01345]" | sed -n 20,40p
for c in 1 2; do echo "== chains $c"; MFM_WS_CHAINS=$c timeout 120 python scripts/lstm_prof.py 2>&1 | head -4; done
