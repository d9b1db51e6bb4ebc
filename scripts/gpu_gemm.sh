#!/bin/bash
# pipelined GEMM: correctness, role timeline, timings (MFM_TCP=0 = register-prefetch kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
timeout 300 python scripts/gemm_trace.py 0 2 2>&1 | grep -E "==|steady|^ +(3|4|22|30) " | tee gpurun_out/gemm_trace.txt
timeout 300 python scripts/gemm_bench.py 1 2>&1 | tee gpurun_out/gemm_bench.txt
