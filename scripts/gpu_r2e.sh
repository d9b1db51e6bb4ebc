#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/step_timeline.py 2>&1 | tee gpurun_out/step_timeline.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5
