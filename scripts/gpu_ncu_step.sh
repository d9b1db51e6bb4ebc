#!/bin/bash
# per-kernel device time of one training step (no graph), serialized by ncu: compare SHARES
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q --tb=short -p no:cacheprovider -x -k "lstm" 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 300 --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu exit $?"
