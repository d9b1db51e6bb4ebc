#!/bin/bash
# ncu --set full of the pipelined GEMM on four representative shapes (second launch of each)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 4 -f -o gpurun_out/gemm_tcp_r1c python scripts/gemm_prof.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_gemm.log; ls -la gpurun_out/*.ncu-rep
