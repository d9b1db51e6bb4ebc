#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_primitives.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_fwd -s 1 -c 1 -o gpurun_out/lstm_fwd python scripts/lstm_prof.py > gpurun_out/ncu_lstm.log 2>&1
echo "ncu exit $?"
bash scripts/gpu_bench_only.sh 2>&1 | head -4
