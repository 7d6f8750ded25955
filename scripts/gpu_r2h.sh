#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest bsvd/conv/cfg3"; timeout 1500 python -m pytest tests/test_bsvd_gpu.py tests/test_conv_gpu.py tests/test_cfg3_gpu.py -m gpu -q --timeout 900 > gpurun_out/pytest_c.log 2>&1; tail -n 5 gpurun_out/pytest_c.log
echo "=== bsvd split profile (twin-tile fast store)"; timeout 300 python scripts/profile_bsvd.py 8 split > gpurun_out/profile_bsvd_split_fast.log 2>&1; head -1 gpurun_out/profile_bsvd_split_fast.log; sed -n 2,5p gpurun_out/profile_bsvd_split_fast.log; sed -n 16,20p gpurun_out/profile_bsvd_split_fast.log
echo "=== bsvd split profile (general epilogue)"; SS4K_NO_SPLIT_FAST=1 timeout 300 python scripts/profile_bsvd.py 8 split > gpurun_out/profile_bsvd_split_nofast.log 2>&1; head -1 gpurun_out/profile_bsvd_split_nofast.log
echo "=== bench cfg3"; timeout 900 python bench.py > gpurun_out/bench_cfg3_h.json 2> gpurun_out/bench_cfg3_h.err; tail -n 1 gpurun_out/bench_cfg3_h.json
