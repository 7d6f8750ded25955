#!/bin/bash
# round 2, second visit: stride-2 streaming convs, fused RDB: parity + bench
mkdir -p gpurun_out
echo "=== pytest conv/bsvd/nets"; timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_bsvd_gpu.py tests/test_nets_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_a.log 2>&1; tail -n 8 gpurun_out/pytest_a.log
echo "=== pytest fullsize"; timeout 1200 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s --timeout 900 > gpurun_out/pytest_full.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_full.log | tail -n 30
echo "=== pytest cfg3"; timeout 900 python -m pytest tests/test_cfg3_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/pytest_cfg3.log 2>&1; grep -v "^\.*$" gpurun_out/pytest_cfg3.log | tail -n 12
echo "=== bench cfg2 fused"; timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_fused.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_fused.log | cut -c1-400
echo "=== bench cfg2 unfused"; SS4K_NO_RDB_FUSE=1 timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_unfused.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_unfused.log | cut -c1-400
echo "=== bench cfg2 fused, counters ignored (WRONG results: cost of the waits)"; SS4K_RDB_DBG=1 timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_fused_nodep.log 2>&1; tail -n 1 gpurun_out/bench_cfg2_fused_nodep.log | cut -c1-400
echo "=== bench cfg3"; timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.log 2>&1; tail -n 1 gpurun_out/bench_cfg3.log | cut -c1-400
