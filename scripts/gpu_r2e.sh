#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?"; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_a.json'))
print(j['ms_per_step'], j['value'], j['e2e']['value'])
for k,v in j['kernels'].items(): print(k, v['ms_per_step'], v['launches'])
PY
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5
