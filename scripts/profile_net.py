"""Per-step times of a net's plan from un-graphed CUDA-event profiling: python scripts/profile_net.py srvgg|rrdb [batch]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L
from ss4k_b200 import realesrgan, engine as E
from oracle import srvgg, rrdbnet

which = sys.argv[1] if len(sys.argv) > 1 else "srvgg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
if which == "srvgg":
    net = srvgg.SRVGGNetCompact(3, 3, 64, 32, 4).eval()
    m = realesrgan.NativeSRVGG(net.state_dict(), num_conv=32, upscale=4, device=0)
else:
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval()
    m = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=23, device=0)
frames = torch.randint(0, 256, (B, 720, 1280, 3), dtype=torch.uint8, device="cuda")
plan = m._plan(B, 720, 1280, L.FMT_U8_NHWC, L.FMT_F16_NCHW)
plan.run(frames); torch.cuda.synchronize()
steps = E.plan_dry(plan.cfg)["steps"]
for _ in range(3):
    prof = plan.profile(frames)
tot = sum(p[0] for p in prof)
print(json.dumps({"net": which, "batch": B, "steps": len(prof), "ms_total": tot, "TFLOP/s": sum(p[1] for p in prof) / tot / 1e9}))
agg = {}
for i, (ms, fl, kd) in enumerate(prof):
    st = steps[i] if i < len(steps) else {}
    key = (kd, st.get("cin"), st.get("cout"), st.get("in_h"), st.get("out_mode"), st.get("act"))
    a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += fl
for k, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"kind {k[0]} cin {k[1]} cout {k[2]} h {k[3]} out_mode {k[4]} act {k[5]}: {n:3d} launches, {ms*1000:9.1f} us total, {ms*1000/n:8.1f} us each, {fl/ms/1e9 if ms else 0:7.0f} TF")
