#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_trace.py 2>&1 | tee gpurun_out/gemm_trace.txt
