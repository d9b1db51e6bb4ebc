#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 8 -o gpurun_out/gemm_tc_r1 python scripts/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_gemm.log; ls -la gpurun_out/*.ncu-rep
