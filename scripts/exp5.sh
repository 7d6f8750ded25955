timeout 600 python -m pytest tests/test_nets_gpu.py -m gpu -q -s --timeout 300 2>&1 | grep -v "^\.*$" | tail -12
timeout 900 python scripts/sweep_configs.py cfg4 cfg5 2>&1 | tail -7
