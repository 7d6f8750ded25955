#!/bin/bash
# DRAM traffic of one frame, ONE ncu pass (two counters of the same unit: no kernel replay, so no save/restore
# of device memory between passes pollutes L2), warm caches as in the real step.
# usage: scripts/gpu_dram.sh <tag> [batch]
TAG=${1:-x}; B=${2:-1}
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
  -k regex:'conv3x3|prep_kernel' -c 420 --csv --log-file gpurun_out/${TAG}_dram_launches_b${B}.csv \
  python bench.py --steps 1 --warmup 1 --batch $B --no-cpu > gpurun_out/${TAG}_dram.log 2>&1
python - <<P
import csv, collections, json
rows = list(csv.reader(open("gpurun_out/${TAG}_dram_launches_b${B}.csv")))
hdr = None; recs = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except ValueError: continue
        recs.setdefault(int(d["ID"]), {"k": d["Kernel Name"]})[d["Metric Name"]] = v
ids = sorted(recs)
# last frame = last 353 launches
last = ids[-353:] if len(ids) >= 353 else ids
rd = sum(recs[i].get("dram__bytes_read.sum", 0) for i in last); wr = sum(recs[i].get("dram__bytes_write.sum", 0) for i in last)
print(json.dumps({"launches": len(last), "dram_read_GB": rd / 1e9, "dram_write_GB": wr / 1e9}))
for i in last[100:110]:
    print(i, recs[i]["k"][:30], round(recs[i].get("dram__bytes_read.sum", 0) / 1e6, 1), round(recs[i].get("dram__bytes_write.sum", 0) / 1e6, 1))
P
