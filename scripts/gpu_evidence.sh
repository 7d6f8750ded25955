#!/bin/bash
# Evidence bundle for profiles/: probe, per-row trace, conv micro-benchmarks, config sweeps, bench lines, launch list.
mkdir -p gpurun_out
timeout 60 ./scripts/mma_issue_probe.bin > gpurun_out/r01_v5_mma_issue_probe.txt 2>&1
timeout 300 python scripts/trace_conv.py > gpurun_out/r01_v5_trace_conv.log 2>&1
timeout 600 python scripts/bench_conv.py > gpurun_out/r01_v5_bench_conv.log 2>&1
timeout 900 python scripts/sweep_configs.py cfg3 cfg4 cfg5 > gpurun_out/r01_v5_sweep_configs.log 2>&1; tail -12 gpurun_out/r01_v5_sweep_configs.log
timeout 600 python bench.py > gpurun_out/r01_v5_bench_b1.json 2>gpurun_out/bench_err.log; tail -c 600 gpurun_out/r01_v5_bench_b1.json
timeout 600 python bench.py --batch 4 --steps 30 > gpurun_out/r01_v5_bench_b4.json 2>>gpurun_out/bench_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 365 --csv --log-file gpurun_out/r01_v5_launches_rrdb720p_b1.csv python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
