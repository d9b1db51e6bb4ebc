#!/bin/bash
# full GPU test suite + bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -15
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_n1.json'))
print(j['ms_per_step'], j['value'], j['e2e']['value'], j['clocks'])
print(json.dumps(j['parity_check']))
print(j['cpu_baseline'])
print(j['roofline']['traffic'], j['roofline']['frac'])
for k,v in list(j['kernels'].items())[:14]: print(k, v['ms_per_step'], v['launches'])
PY
