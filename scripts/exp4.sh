timeout 600 python -m pytest tests/test_bsvd_gpu.py tests/test_conv_gpu.py tests/test_colour_gpu.py tests/test_nets_gpu.py -m gpu -q -s --timeout 300 2>&1 | grep -v "^\.*$" | tail -22
timeout 600 python scripts/sweep_configs.py cfg3 2>&1 | tail -3
timeout 300 python bench.py --steps 50 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['clocks']); print(d['hbm_kernels'])"
