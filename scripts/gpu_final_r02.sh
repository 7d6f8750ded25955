#!/bin/bash
# Final evidence of round 2 (tag r02_p): whole GPU suite, bench lines, BSVD profiles, streaming, launch list + ncu captures.
mkdir -p gpurun_out
T=r02_p
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/${T}_pytest_gpu.log 2>&1; grep -v "^\.*$" gpurun_out/${T}_pytest_gpu.log | tail -n 12
echo "=== bench default (cfg3 split)"; timeout 1200 python bench.py > gpurun_out/${T}_bench_cfg3_n1.json 2> gpurun_out/${T}_bench_cfg3_n1.err; tail -n 1 gpurun_out/${T}_bench_cfg3_n1.json | cut -c1-300
echo "=== bench cfg3 f16 bsvd"; timeout 900 python bench.py --bsvd f16 --no-cpu > gpurun_out/${T}_bench_cfg3_f16_n1.json 2>/dev/null; tail -n 1 gpurun_out/${T}_bench_cfg3_f16_n1.json | cut -c1-200
echo "=== bench cfg2"; timeout 600 python bench.py --workload cfg2 --no-cpu > gpurun_out/${T}_bench_cfg2_n1.json 2>/dev/null; tail -n 1 gpurun_out/${T}_bench_cfg2_n1.json | cut -c1-200
echo "=== bsvd profiles"; timeout 300 python scripts/profile_bsvd.py 8 split nv12 > gpurun_out/${T}_profile_bsvd_split_nv12.log 2>&1; head -1 gpurun_out/${T}_profile_bsvd_split_nv12.log
timeout 300 python scripts/profile_bsvd.py 8 f16 nv12 > gpurun_out/${T}_profile_bsvd_f16_nv12.log 2>&1; head -1 gpurun_out/${T}_profile_bsvd_f16_nv12.log
echo "=== bsvd streaming"; timeout 600 python scripts/bench_bsvd_stream.py > gpurun_out/${T}_bench_bsvd_stream.log 2>&1; cat gpurun_out/${T}_bench_bsvd_stream.log | tail -5
echo "=== role traces"; timeout 300 python scripts/bench_bsvd_fullres.py > gpurun_out/${T}_bench_bsvd_fullres.log 2>&1; grep -c src gpurun_out/${T}_bench_bsvd_fullres.log
echo "=== ncu"
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_cfg3.csv python bench.py --steps 1 --warmup 1 --clip 2 --no-cpu > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-120
timeout 900 $NCU -k regex:conv3x3_stream -s 32 -c 6 -f -o gpurun_out/${T}_bsvd_split_full python scripts/profile_bsvd.py 8 split nv12 > gpurun_out/ncu_a.log 2>&1
timeout 900 $NCU -k regex:conv3x3_stream -s 390 -c 5 -f -o gpurun_out/${T}_trunk_full python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_b.log 2>&1
for r in bsvd_split trunk; do
  ncu -i gpurun_out/${T}_${r}_full.ncu-rep --page raw --csv > gpurun_out/${T}_${r}_full_raw.csv 2>/dev/null
  wc -l gpurun_out/${T}_${r}_full_raw.csv
done
rm -f gpurun_out/*.ncu-rep
ls gpurun_out | grep ${T}
