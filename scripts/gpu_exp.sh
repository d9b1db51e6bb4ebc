#!/bin/bash
# A/B of one engine switch (timeline tail + step time); usage: gpu_exp.sh ENVVAR
for v in 1 0; do
  echo "== $1=$v"
  env $1=$v timeout 300 python scripts/step_timeline.py 2>&1 | grep -E "bwd:lstm|bwd:join wgrads|bwd:att1"
  env $1=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'samples/s')"
done
