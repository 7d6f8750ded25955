mkdir -p gpurun_out
for b in 1 2 4; do timeout 600 python bench.py --batch $b --no-cpu --steps $((400/b)) --warmup 5 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('b',d['config']['frames_per_step_per_gpu'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['clocks'],'whole',round(d['roofline']['whole_step_tflops']),'kern',round(d['roofline']['achieved']))"; done
