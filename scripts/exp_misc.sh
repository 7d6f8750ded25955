#!/bin/bash
run() { # label, env..., args
  local label=$1; shift
  env "$@" | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label fps', round(d['value'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
}
run "f16 b1 discard0" SS4K_DISCARD=0 timeout 200 python bench.py --no-cpu --steps 150 2>/dev/null
run "f16 b1 discard1" SS4K_DISCARD=1 timeout 200 python bench.py --no-cpu --steps 150 2>/dev/null
run "f16 b2 discard1" SS4K_DISCARD=1 timeout 200 python bench.py --no-cpu --steps 80 --batch 2 2>/dev/null
run "bf16 b1" SS4K_DISCARD=0 timeout 200 python bench.py --no-cpu --steps 150 --dtype bf16 2>/dev/null
run "bf16 b2" SS4K_DISCARD=0 timeout 200 python bench.py --no-cpu --steps 80 --dtype bf16 --batch 2 2>/dev/null
run "bf16 b2 discard1" SS4K_DISCARD=1 timeout 200 python bench.py --no-cpu --steps 80 --dtype bf16 --batch 2 2>/dev/null
echo "dram discard1"; SS4K_DISCARD=1 bash scripts/gpu_dram.sh s4_disc 1 2>&1 | tail -11 | head -7
SS4K_DISCARD=1 timeout 300 python -m pytest tests/test_fullsize_gpu.py tests/test_nets_gpu.py -m gpu -q -x 2>&1 | tail -2
