#!/bin/bash
# pipelined GEMM: correctness, role timeline, timings (MFM_TCP=0 = register-prefetch kernel, MFM_TCP_L2 = L2 promotion bytes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
for l2 in 128 256 0; do
MFM_TCP_L2=$l2 timeout 300 python scripts/gemm_trace.py 0 2 2>&1 | grep -E "==|steady" | sed "s/^/[L2=$l2] /" | tee -a gpurun_out/gemm_trace.txt
MFM_TCP_L2=$l2 timeout 300 python scripts/gemm_bench.py 1 2>&1 | grep -E "att1_fc1|att1_fc2|dAtt|dW11|dW_ih" | sed "s/^/[L2=$l2] /" | tee -a gpurun_out/gemm_bench.txt
done
