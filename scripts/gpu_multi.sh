#!/bin/bash
# multi-GPU bench lines (torchrun, one rank per GPU): usage gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_cfg3_n$N.log 2>gpurun_out/bench_cfg3_n$N.err; tail -n 1 gpurun_out/bench_cfg3_n$N.log | cut -c1-700; tail -n 3 gpurun_out/bench_cfg3_n$N.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg2 --steps 100 --warmup 5 > gpurun_out/bench_cfg2_n$N.log 2>gpurun_out/bench_cfg2_n$N.err; tail -n 1 gpurun_out/bench_cfg2_n$N.log | cut -c1-400
