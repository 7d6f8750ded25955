#!/bin/bash
# First GPU visit: descriptor-mode diagnostics, parity tests, a first bench line, a launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== diag" ; timeout 900 python scripts/diag_conv.py > gpurun_out/diag.log 2>&1 ; tail -c 6000 gpurun_out/diag.log
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1 ; tail -n 40 gpurun_out/pytest_gpu.log
echo "=== pytest gpu (no -x, nets)" ; timeout 1200 python -m pytest tests/test_nets_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/pytest_nets.log 2>&1 ; tail -n 30 gpurun_out/pytest_nets.log
echo "=== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1 ; tail -n 5 gpurun_out/bench.log
if grep -q '"metric"' gpurun_out/bench.log; then
  echo "=== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
  tail -n 3 gpurun_out/ncu_bench.log
fi
