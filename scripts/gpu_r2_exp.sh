#!/bin/bash
# parity, then the step time with and without the weight-gradient GEMMs (experiment: what do they cost end to end?)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
for v in "" 1; do
  MFM_SKIP_WGRAD=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('skip_wgrad=[$v]', round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'samples/s; e2e', round(j['e2e']['ms_per_step'],3))"
done
