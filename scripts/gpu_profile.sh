#!/bin/bash
# bench line + ncu launch list + ncu --set full capture of the dominant kernel (one RDB: conv1..conv5)
mkdir -p gpurun_out
echo "=== bench b${B:-1}" ; timeout 600 python bench.py --steps 20 --warmup 3 --batch ${B:-1} > gpurun_out/bench_prof.log 2>&1 ; tail -n 1 gpurun_out/bench_prof.log | cut -c1-600
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch ${B:-1} --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
echo "=== ncu full (6 stream-kernel launches inside the trunk)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_stream -s 30 -c 6 -o gpurun_out/prof_stream -f python bench.py --steps 1 --warmup 0 --batch ${B:-1} --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/
