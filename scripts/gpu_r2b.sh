#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 12 -c 4 -f -o gpurun_out/lstm_ws_r2b python scripts/lstm_prof.py > gpurun_out/ncu_lstm.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_lstm.log
