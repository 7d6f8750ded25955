#!/bin/bash
# Evidence bundle (tag T) for profiles/: bench lines, launch list, ncu full-set captures (raw CSV pages), clean
# single-pass DRAM traffic, per-row trace, conv micro-benchmarks.
mkdir -p gpurun_out
T=r01_v7
timeout 600 python bench.py > gpurun_out/${T}_bench_b1.json 2>gpurun_out/${T}_bench_err.log; tail -c 400 gpurun_out/${T}_bench_b1.json; echo
timeout 300 python bench.py --batch 4 --steps 30 --no-cpu > gpurun_out/${T}_bench_b4.json 2>>gpurun_out/${T}_bench_err.log
timeout 300 python scripts/trace_conv.py > gpurun_out/${T}_trace_conv.log 2>&1
timeout 600 python scripts/bench_conv.py > gpurun_out/${T}_bench_conv.log 2>&1
timeout 200 python scripts/trace_epilogue.py > gpurun_out/${T}_trace_epilogue.log 2>&1
timeout 100 python scripts/profile_net.py rrdb 1 > gpurun_out/${T}_profile_rrdb.log 2>&1
timeout 100 python scripts/profile_net.py srvgg 4 > gpurun_out/${T}_profile_srvgg.log 2>&1
timeout 100 python scripts/profile_bsvd.py 8 > gpurun_out/${T}_profile_bsvd.log 2>&1
timeout 800 python scripts/sweep_configs.py cfg3 cfg4 cfg5 live > gpurun_out/${T}_sweep_configs.log 2>&1
bash scripts/gpu_dram.sh ${T} 1 2>&1 | tail -12 | head -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 365 --csv --log-file gpurun_out/${T}_launches_rrdb720p_b1.csv python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_stream -s 36 -c 6 -f -o gpurun_out/${T}_stream_full python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_stream -s 347 -c 5 -f -o gpurun_out/${T}_tail_full python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_tail.log 2>&1
for r in stream tail; do
  ncu -i gpurun_out/${T}_${r}_full.ncu-rep --page raw --csv > gpurun_out/${T}_${r}_full_raw.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -20
