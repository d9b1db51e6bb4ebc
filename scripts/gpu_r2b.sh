#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws_fwd -s 1 -c 1 -f -o gpurun_out/lstm_ws_r2c python scripts/lstm_prof.py > gpurun_out/ncu_lstm.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_lstm.log
