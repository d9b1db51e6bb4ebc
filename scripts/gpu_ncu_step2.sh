#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 225 --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu exit $?"
