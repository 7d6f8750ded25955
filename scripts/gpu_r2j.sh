#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest bsvd/cfg3/colour"; timeout 1500 python -m pytest tests/test_bsvd_gpu.py tests/test_cfg3_gpu.py tests/test_colour_gpu.py tests/test_multi_gpu_gpu.py tests/test_service_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_j.log 2>&1; tail -n 15 gpurun_out/pytest_j.log
echo "=== bsvd split profile, NV12 frames"; timeout 300 python scripts/profile_bsvd.py 8 split nv12 > gpurun_out/profile_bsvd_split_j.log 2>&1; head -4 gpurun_out/profile_bsvd_split_j.log
echo "=== bench cfg3"; timeout 900 python bench.py > gpurun_out/bench_cfg3_j.json 2> gpurun_out/bench_cfg3_j.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_cfg3_j.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['config']['ms_per_frame'],d['gpu_launches'])"
