#!/bin/bash
# round evidence: launch list of the whole bench run (one step is cut out of it afterwards), ncu --set full of the
# pipelined GEMM on four representative shapes, the step timeline, the GEMM role trace and the bench line itself
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 4 -f -o gpurun_out/gemm_tcp_r1g python scripts/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu full exit $?"
timeout 300 python scripts/step_timeline.py > gpurun_out/step_timeline.txt 2>&1; echo "timeline exit $?"
timeout 300 python scripts/gemm_trace.py > gpurun_out/gemm_trace.txt 2>&1; echo "trace exit $?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
