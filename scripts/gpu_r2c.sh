#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q --tb=short -p no:cacheprovider -k lstm 2>&1 | tail -8
timeout 300 python scripts/lstm_prof.py 2>&1 | tee gpurun_out/lstm_prof.txt
