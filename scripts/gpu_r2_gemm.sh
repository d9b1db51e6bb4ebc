#!/bin/bash
# pipelined GEMM: correctness first, then timings for both K-chunk sizes (MFM_TCP=0 = register-prefetch kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
for bk in 32 16; do
  MFM_TCP_BK=$bk timeout 300 python scripts/gemm_bench.py 1 2>&1 | tee gpurun_out/gemm_bench_bk$bk.txt
done
