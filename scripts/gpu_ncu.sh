#!/bin/bash
# ncu evidence for the round: (1) full-set capture of consecutive trunk launches of the streaming conv kernel
# (+ one <64> and the <16> tail kernel), (2) launch list of one frame (gpu__time_duration only).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_stream -s 36 -c 6 -f -o gpurun_out/r01_v5_stream_full python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_stream -s 347 -c 5 -f -o gpurun_out/r01_v5_tail_full python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_tail.log 2>&1
tail -2 gpurun_out/ncu_tail.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 353 --csv --log-file gpurun_out/r01_v5_launches_rrdb720p_b1.csv python bench.py --steps 1 --warmup 0 --batch 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-200
ls -la gpurun_out/
