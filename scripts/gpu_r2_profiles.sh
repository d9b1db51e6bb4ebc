#!/bin/bash
# round-2 evidence: bench line, launch list of one eager step, ncu --set full of the GEMM shapes / recurrence kernels /
# memory recurrence, step timeline
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_ps_kernel|gemm_tcp_kernel" -s 4 -c 4 -f -o gpurun_out/gemm_tcp_r2 python scripts/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -c 8 -f -o gpurun_out/lstm_ws_r2 env T=20 python scripts/lstm_prof_once.py > gpurun_out/ncu_lstm.log 2>&1; echo "ncu lstm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mem_ws -s 4 -c 2 -f -o gpurun_out/mem_ws_r2 env T=20 python scripts/mem_prof.py > gpurun_out/ncu_mem.log 2>&1; echo "ncu mem exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/step_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity-check > gpurun_out/ncu_step.log 2>&1; echo "ncu launches exit $?"
timeout 300 python scripts/step_timeline.py > gpurun_out/step_timeline.txt 2>&1; echo "timeline exit $?"; cat gpurun_out/step_timeline.txt
