#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_a.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_a.log; tail -12 gpurun_out/pytest_a.log
timeout 300 python scripts/gemm_bench.py 2>&1 | grep "path 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 8 -o gpurun_out/gemm_tc_r1v2 python scripts/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"
bash scripts/gpu_bench_only.sh
