#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench exit $?"; cat gpurun_out/gemm_bench.log | tail -30
