#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/step_timeline.py 2>&1 | tee gpurun_out/step_timeline.txt
