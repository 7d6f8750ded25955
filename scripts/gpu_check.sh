#!/bin/bash
# Routine GPU visit: parity tests, bench line (default), launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/pytest_gpu.log 2>&1 ; grep -v "^\.*$" gpurun_out/pytest_gpu.log | tail -n 45
echo "=== bench default" ; timeout 900 python bench.py > gpurun_out/bench.log 2>&1 ; tail -n 3 gpurun_out/bench.log
