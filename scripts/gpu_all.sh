#!/bin/bash
# full GPU test suite + step timeline + step bench with per-GEMM-shape timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -6
timeout 300 python scripts/step_timeline.py 2>&1 | tee gpurun_out/step_timeline.txt
bash scripts/gpu_bench_only.sh 2>&1 | head -80
