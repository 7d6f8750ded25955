#!/bin/bash
# L2 eviction-priority hint variants of the RRDB dense block (SS4K_L2_HINTS: conv1-4 loads, conv1-4 stores, conv5 loads, conv5 stores)
for h in 0000 1121 1111 1101 1120 0020 1100; do
  SS4K_L2_HINTS=$h timeout 200 python bench.py --no-cpu --steps 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hints $h fps', round(d['value'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
done
for h in 1121 1111; do
  echo "dram hints $h"; SS4K_L2_HINTS=$h bash scripts/gpu_dram.sh s4_h$h 1 2>&1 | tail -11 | head -7
done
