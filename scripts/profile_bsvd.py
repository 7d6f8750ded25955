"""Per-step times of the BSVD-32 clip plan (F frames of 1280x720) from un-graphed CUDA-event profiling."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L
from ss4k_b200 import bsvd as nbsvd, engine as E
from oracle import bsvd as obsvd

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SPLIT = len(sys.argv) > 2 and sys.argv[2] == "split"
sd = obsvd.build_bsvd32(0, weight_scale=1.0 if SPLIT else 0.5)
den = nbsvd.NativeBSVD(sd, device=0, act_mode=L.ACT_F16_SPLIT if SPLIT else L.ACT_F16, out_dtype=torch.float16)
NV12 = len(sys.argv) > 3 and sys.argv[3] == "nv12"   # frame-format entry: the first conv's loader decodes the frames
if NV12:
    x = torch.randint(16, 236, (F, 1080, 1280), dtype=torch.uint8, device="cuda")
    y = den.denoise_frames(x, 720, 1280, 0.075, nv12=True); torch.cuda.synchronize()
else:
    x = torch.rand(1, F, 4, 720, 1280, device="cuda")
    y = den(x); torch.cuda.synchronize()
plan = list(den._plans._d.values())[0]
cfg = plan.cfg
dry = E.plan_dry(cfg)
steps = dry["steps"]
xin = x if NV12 else (x[0].contiguous() if x.dtype == torch.float32 else x[0].float().contiguous())
prof = None
for _ in range(3):
    prof = plan.profile(xin)
tot = sum(p[0] for p in prof)
print(json.dumps({"frames": F, "steps": len(prof), "ms_total": tot, "ms_per_frame": tot / F, "TFLOP/s": sum(p[1] for p in prof) / tot / 1e9}))
names = [s.get("name", s.get("kind")) for s in steps]
for i, (ms, fl, kd) in enumerate(prof):
    st = steps[i] if i < len(steps) else {}
    print(f"{i:3d} kind {kd} {ms*1000:8.1f} us  {fl/ms/1e9 if ms > 0 else 0:7.0f} TF  {st.get('name','prep')}  cin {st.get('cin')} cout {st.get('cout')} {st.get('in_h')}x{st.get('in_w')} mode {st.get('mode')} split {st.get('split')}")
