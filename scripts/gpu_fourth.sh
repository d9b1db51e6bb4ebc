#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_a.log 2>&1
echo "pytest A exit $?" >> gpurun_out/pytest_a.log; tail -25 gpurun_out/pytest_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --durations=5 > gpurun_out/pytest_b.log 2>&1
echo "pytest B exit $?" >> gpurun_out/pytest_b.log; tail -25 gpurun_out/pytest_b.log
timeout 300 python scripts/gemm_bench.py 2>&1 | grep "path 1" 
bash scripts/gpu_bench_only.sh
