#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
echo "=== pytest multi_gpu/service run $i"; timeout 900 python -m pytest tests/test_multi_gpu_gpu.py tests/test_service_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_k$i.log 2>&1; tail -n 3 gpurun_out/pytest_k$i.log
done
grep -h "self-probe" gpurun_out/pytest_k*.log | head -5
