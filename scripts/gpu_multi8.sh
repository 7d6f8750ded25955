#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_cfg3_n$N.log 2>gpurun_out/bench_cfg3_n$N.err; tail -n 1 gpurun_out/bench_cfg3_n$N.log | cut -c1-600; tail -n 3 gpurun_out/bench_cfg3_n$N.err | cut -c1-300
