#!/bin/bash
# round 2: fused RDB with weight ring: parity, trace, bench
mkdir -p gpurun_out
echo "=== pytest fullsize"; timeout 1200 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s --timeout 900 -k "fused or full_frame_vs" > gpurun_out/pytest_full.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_full.log | tail -n 12
echo "=== trace"; timeout 300 python scripts/trace_rdb.py > gpurun_out/trace_rdb.log 2>&1; tail -n 40 gpurun_out/trace_rdb.log
echo "=== bench cfg2 fused"; timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_fused.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_fused.log | cut -c1-300
echo "=== bench cfg2 fused, counters ignored"; SS4K_RDB_DBG=1 timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_fused_nodep.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_fused_nodep.log | cut -c1-300
echo "=== bench cfg2 unfused"; SS4K_NO_RDB_FUSE=1 timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_unfused.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_unfused.log | cut -c1-300
