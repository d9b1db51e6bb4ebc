#!/bin/bash
# which part of the pipelined GEMM bounds it?  MFM_TC_DEBUG: 2 skip convert, 4 skip epilogue, 8 skip loads
mkdir -p gpurun_out
for dbg in 0 8 4 2 6 12 14; do
  MFM_TC_DEBUG=$dbg MFM_TCP_BK=32 timeout 300 python scripts/gemm_bench.py 1 2>&1 | grep -E "att1_fc1|att1_fc2|dW11|dAtt" | sed "s/^/[dbg=$dbg] /" | tee -a gpurun_out/gemm_bench_dbg.txt
done
