#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/gemm_tc_nn python scripts/gemm_prof2.py > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu exit $?"
