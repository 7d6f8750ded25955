#!/bin/bash
# parity tests + micro benchmark + bench lines (batch 1 / 4, PDL on / off)
mkdir -p gpurun_out
echo "=== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1 ; tail -n 8 gpurun_out/pytest_gpu.log
echo "=== micro bench" ; timeout 300 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1 ; tail -n 3 gpurun_out/bench_conv.log
echo "=== trace" ; timeout 300 python scripts/trace_conv.py > gpurun_out/trace_conv.log 2>&1 ; tail -n 2 gpurun_out/trace_conv.log
echo "=== bench b1" ; timeout 600 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu > gpurun_out/bench_b1.log 2>&1 ; tail -n 1 gpurun_out/bench_b1.log | cut -c1-300
echo "=== bench b1 no PDL" ; SS4K_NO_PDL=1 timeout 600 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu > gpurun_out/bench_b1_nopdl.log 2>&1 ; tail -n 1 gpurun_out/bench_b1_nopdl.log | cut -c1-300
echo "=== bench b4" ; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1 ; tail -n 1 gpurun_out/bench.log | cut -c1-300
