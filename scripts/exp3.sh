mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
for e in "A=1" "SS4K_NO_W_PREFETCH=1" "A=1" "SS4K_NO_W_PREFETCH=1"; do
env $e timeout 600 python bench.py --batch 1 --no-cpu --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$e b',d['config']['frames_per_step_per_gpu'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['clocks'],'whole',round(d['roofline']['whole_step_tflops']),'kern',round(d['roofline']['achieved']))"
done
