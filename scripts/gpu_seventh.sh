#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -4
bash scripts/gpu_bench_only.sh 2>&1 | head -16
