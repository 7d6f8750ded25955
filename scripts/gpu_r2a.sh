#!/bin/bash
# round 2, first visit: cfg3 parity test + cfg3 / cfg2 bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "=== pytest cfg3"; timeout 900 python -m pytest tests/test_cfg3_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/pytest_cfg3.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_cfg3.log | tail -n 30
echo "=== bench cfg3"; timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_cfg3.log 2>&1; tail -n 3 gpurun_out/bench_cfg3.log | cut -c1-3000
echo "=== bench cfg3 f16"; timeout 600 python bench.py --steps 3 --warmup 3 --bsvd f16 --no-cpu > gpurun_out/bench_cfg3_f16.log 2>&1; tail -n 1 gpurun_out/bench_cfg3_f16.log | cut -c1-600
echo "=== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2.log 2>&1; tail -n 1 gpurun_out/bench_cfg2.log | cut -c1-600
echo "=== bsvd split profile"; timeout 300 python scripts/profile_bsvd.py 8 split > gpurun_out/profile_bsvd_split.log 2>&1; head -40 gpurun_out/profile_bsvd_split.log
