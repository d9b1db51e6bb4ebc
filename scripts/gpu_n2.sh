#!/bin/bash
# 2-GPU weak-scaling check of the bench contract (one rank per GPU over NCCL)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 exit $?"; cut -c1-700 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
