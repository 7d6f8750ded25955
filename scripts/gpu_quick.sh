#!/bin/bash
# quick GPU iteration: conv + net parity tests, per-row trace, bench b=1 steady state
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -5
timeout 300 python scripts/trace_conv.py > gpurun_out/trace.log 2>&1; grep -A12 '"flags": 0' gpurun_out/trace.log | head -60
timeout 600 python bench.py --batch 1 --no-cpu --steps 200 --warmup 5 > gpurun_out/bench_b1.log 2>&1; tail -1 gpurun_out/bench_b1.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('b',d['config']['frames_per_step_per_gpu'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['clocks'],'whole',round(d['roofline']['whole_step_tflops']),'kern',round(d['roofline']['achieved']))"
