#!/bin/bash
# round 2: scout warp in the streaming conv kernel: parity + micro benchmarks + bench
mkdir -p gpurun_out
echo "=== pytest conv/bsvd/nets"; timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_bsvd_gpu.py tests/test_nets_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_a.log 2>&1; tail -n 6 gpurun_out/pytest_a.log
echo "=== bench_conv"; timeout 600 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1; grep -E '"flags": (0|7),' gpurun_out/bench_conv.log | head -12 | cut -c1-150
echo "=== trace_conv"; timeout 300 python scripts/trace_conv.py > gpurun_out/trace_conv.log 2>&1; head -12 gpurun_out/trace_conv.log | cut -c1-200
echo "=== bench cfg2 unfused"; SS4K_NO_RDB_FUSE=1 timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_unfused.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_unfused.log | cut -c1-300
echo "=== bench cfg2 fused"; timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_fused.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_fused.log | cut -c1-300
echo "=== pytest fullsize"; timeout 1200 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s --timeout 900 > gpurun_out/pytest_full.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_full.log | tail -n 12
