#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest bsvd/conv/nets/cfg3"; timeout 1500 python -m pytest tests/test_bsvd_gpu.py tests/test_conv_gpu.py tests/test_nets_gpu.py tests/test_cfg3_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_n.log 2>&1; tail -n 4 gpurun_out/pytest_n.log
echo "=== bsvd split (no stage for non-TMA stores)"; timeout 300 python scripts/profile_bsvd.py 8 split nv12 > gpurun_out/profile_bsvd_split_n.log 2>&1; cat gpurun_out/profile_bsvd_split_n.log | cut -c1-110
echo "=== bsvd split (stage kept)"; SS4K_NO_PS2_FAST=1 timeout 300 python scripts/profile_bsvd.py 8 split nv12 > gpurun_out/profile_bsvd_split_n_keep.log 2>&1; head -1 gpurun_out/profile_bsvd_split_n_keep.log
echo "=== bsvd f16"; timeout 300 python scripts/profile_bsvd.py 8 f16 nv12 > gpurun_out/profile_bsvd_f16_n.log 2>&1; head -1 gpurun_out/profile_bsvd_f16_n.log
SS4K_NO_PS2_FAST=1 timeout 300 python scripts/profile_bsvd.py 8 f16 nv12 > gpurun_out/profile_bsvd_f16_n_keep.log 2>&1; head -1 gpurun_out/profile_bsvd_f16_n_keep.log
echo "=== cfg2"; timeout 600 python bench.py --workload cfg2 --no-cpu > gpurun_out/bench_cfg2_n.json 2>/dev/null; tail -n 1 gpurun_out/bench_cfg2_n.json | cut -c1-160
SS4K_KEEP_STAGE=1 timeout 600 python bench.py --workload cfg2 --no-cpu > gpurun_out/bench_cfg2_n_keep.json 2>/dev/null; tail -n 1 gpurun_out/bench_cfg2_n_keep.json | cut -c1-160
