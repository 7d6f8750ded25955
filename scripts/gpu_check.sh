#!/bin/bash
# Routine GPU visit: parity tests, bench line (b=1 and default), launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1 ; tail -n 25 gpurun_out/pytest_gpu.log
echo "=== bench default" ; timeout 900 python bench.py > gpurun_out/bench.log 2>&1 ; tail -n 3 gpurun_out/bench.log
echo "=== bench b1" ; timeout 900 python bench.py --batch 1 --no-cpu > gpurun_out/bench_b1.log 2>&1 ; tail -n 3 gpurun_out/bench_b1.log
