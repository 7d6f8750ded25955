#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest cfg3 / bsvd / nets"; timeout 1500 python -m pytest tests/test_cfg3_gpu.py tests/test_bsvd_gpu.py tests/test_nets_gpu.py tests/test_colour_gpu.py tests/test_service_gpu.py -m gpu -q -s --timeout 900 > gpurun_out/pytest_b.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_b.log | tail -n 25
N=2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_cfg3_n$N.log 2>gpurun_out/bench_cfg3_n$N.err; tail -n 1 gpurun_out/bench_cfg3_n$N.log | cut -c1-400; tail -n 3 gpurun_out/bench_cfg3_n$N.err | cut -c1-300
