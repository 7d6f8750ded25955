#!/bin/bash
mkdir -p gpurun_out
echo "=== micro bench" ; timeout 600 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1 ; cat gpurun_out/bench_conv.log | tail -n 60
echo "=== ncu full on one RDB (convs 18..23 of the plan)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 18 -c 5 -o gpurun_out/prof_rdb -f python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_full.log; ls -la gpurun_out/
