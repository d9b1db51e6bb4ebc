#!/bin/bash
for d in 0 15; do MFM_TC_DEBUG=$d timeout 120 python scripts/gemm_floor.py 2>&1 | tail -8; done
