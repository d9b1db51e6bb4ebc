#!/bin/bash
mkdir -p gpurun_out
MFM_BENCH_GEMM_SHAPES=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_shapes.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_shapes.json'))
print({k:j[k] for k in ('value','ms_per_step','launches_per_step')})
for k,v in j['kernels'].items(): print("  %-22s %s"%(k,v))
for k,v in j.get('gemm_shapes_ms',{}).items(): print("  %-34s %8.4f ms x%d"%(k,v[0],v[1]))
PY
tail -5 gpurun_out/bench.err
