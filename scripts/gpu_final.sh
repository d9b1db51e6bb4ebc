#!/bin/bash
# end-of-round check: full GPU suite, smoke, launch list, timeline, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"
timeout 300 python scripts/step_timeline.py > gpurun_out/step_timeline.txt 2>&1; echo "timeline exit $?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_n1.json
