#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "trainer" 2>&1 | tail -3
bash scripts/gpu_ncu_step2.sh
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench.json')); print({k:j[k] for k in ('value','ms_per_step','launches_per_step','clocks')}); print(j['e2e']); print(j['roofline']); print(j.get('cpu_baseline'))"
