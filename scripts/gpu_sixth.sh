#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python scripts/gemm_bench.py 2>&1 | grep "path 1"
bash scripts/gpu_bench_only.sh 2>&1 | head -14
