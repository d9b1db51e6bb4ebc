#!/bin/bash
# round 2, first pass: new recurrence kernels -- primitive + parity tests, stand-alone timing, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_primitives.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25
timeout 300 python scripts/lstm_prof.py 2>&1 | tee gpurun_out/lstm_prof.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm_tc.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -25
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?"; cut -c1-900 gpurun_out/bench_a.json
