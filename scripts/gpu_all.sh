#!/bin/bash
# Whole GPU visit: every GPU test, then the bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_gpu.log | tail -n 40
echo "=== bench default (cfg3)"; timeout 1200 python bench.py > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-600
echo "=== bench cfg2"; timeout 600 python bench.py --workload cfg2 --no-cpu > gpurun_out/bench_cfg2.log 2>&1; tail -n 1 gpurun_out/bench_cfg2.log | cut -c1-300
