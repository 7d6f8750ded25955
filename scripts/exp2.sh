mkdir -p gpurun_out
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -c 1100 --csv --log-file gpurun_out/dram_b1.csv python bench.py --steps 1 --warmup 2 --batch 1 --no-cpu > gpurun_out/ncu_dram.log 2>&1
tail -2 gpurun_out/ncu_dram.log | cut -c1-300
