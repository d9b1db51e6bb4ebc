#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"; python -c "
import json; j=json.load(open('gpurun_out/bench_n2.json')); print({k:j[k] for k in ('value','ms_per_step','n_gpus','launches_per_step')}, j['e2e'])"; tail -3 gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 exit $?"; python -c "
import json; j=json.load(open('gpurun_out/bench_n1.json')); print({k:j[k] for k in ('value','ms_per_step','n_gpus','launches_per_step')}, j['e2e'], j.get('cpu_baseline'))"
