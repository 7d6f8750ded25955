#!/bin/bash
# ncu --set full of one residual dense block of the trunk WITHOUT the cache flush between replay passes: the dense block's
# working set (118 MB per frame) lives in the 126 MB L2 in the real step, which a cold-cache capture does not show
mkdir -p gpurun_out
T=r02_q
timeout 900 ncu --set full --cache-control none --clock-control none -k regex:conv3x3_stream -s 390 -c 5 -f -o gpurun_out/${T}_trunk_warm python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_w.log 2>&1
ncu -i gpurun_out/${T}_trunk_warm.ncu-rep --page raw --csv > gpurun_out/${T}_trunk_warm_raw.csv 2>/dev/null
wc -l gpurun_out/${T}_trunk_warm_raw.csv
rm -f gpurun_out/*.ncu-rep
