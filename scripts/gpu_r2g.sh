#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest bsvd/conv/cfg3"; timeout 1500 python -m pytest tests/test_bsvd_gpu.py tests/test_conv_gpu.py tests/test_cfg3_gpu.py -m gpu -q --timeout 900 > gpurun_out/pytest_c.log 2>&1; tail -n 5 gpurun_out/pytest_c.log
echo "=== bsvd split profile (min slots 3)"; timeout 300 python scripts/profile_bsvd.py 8 split > gpurun_out/profile_bsvd_split_s3.log 2>&1; head -1 gpurun_out/profile_bsvd_split_s3.log; sed -n 8,12p gpurun_out/profile_bsvd_split_s3.log
echo "=== bsvd split profile (min slots 4)"; SS4K_MIN_SLOTS=4 timeout 300 python scripts/profile_bsvd.py 8 split > gpurun_out/profile_bsvd_split_s4.log 2>&1; head -1 gpurun_out/profile_bsvd_split_s4.log; sed -n 8,12p gpurun_out/profile_bsvd_split_s4.log
echo "=== bsvd split profile (min slots 6)"; SS4K_MIN_SLOTS=6 timeout 300 python scripts/profile_bsvd.py 8 split > gpurun_out/profile_bsvd_split_s6.log 2>&1; head -1 gpurun_out/profile_bsvd_split_s6.log; sed -n 8,12p gpurun_out/profile_bsvd_split_s6.log
