#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench.json'))
print({k:j[k] for k in ('value','ms_per_step','launches_per_step','clocks')}); print(j['e2e']); print(j.get('cpu_baseline'))
for k,v in j['kernels'].items(): print("  %-22s %s"%(k,v))
PY
tail -5 gpurun_out/bench.err
