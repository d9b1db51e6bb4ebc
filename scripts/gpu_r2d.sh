#!/bin/bash
timeout 120 python scripts/lstm_trace.py 2>&1 | tee gpurun_out/lstm_trace.txt | head -120
