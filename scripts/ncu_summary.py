"""Condenses `ncu --page raw --csv` output into one line per metric (columns = captured launches)."""
import csv, sys
KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
if hdr is None:
    sys.exit("no ncu table in " + sys.argv[1])
names, units = rows[hdr], rows[hdr + 1]
data = [r for r in rows[hdr + 2:] if len(r) == len(names)]
print(sys.argv[2] if len(sys.argv) > 2 else "")
for k in KEEP:
    if k not in names:
        continue
    i = names.index(k)
    vals = [r[i][:40] for r in data]
    print(f"{k} [{units[i]}]: " + " | ".join(vals))
