#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest bsvd/service"; timeout 1500 python -m pytest tests/test_bsvd_gpu.py tests/test_service_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_i.log 2>&1; tail -n 15 gpurun_out/pytest_i.log
