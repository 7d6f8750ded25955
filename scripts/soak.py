"""Soak run: many repetitions of the headline pipeline and of the ring-buffer stream with result checks (every repetition of the
same input must reproduce the first result bit for bit; a stuck barrier would trip the kernels' watchdog trap)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L
from ss4k_b200 import bsvd as nb, realesrgan
from ss4k_b200.pipeline import DenoiseUpscalePipeline
from oracle import bsvd as ob, rrdbnet

torch.manual_seed(0)
rr = rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval()
sr = realesrgan.NativeRRDBNet(rr.state_dict(), scale=2, num_block=23, device=0)
den = nb.NativeBSVD(ob.build_bsvd32(0), device=0, act_mode=L.ACT_F16_SPLIT, out_dtype=torch.float16)
pipe = DenoiseUpscalePipeline(den, sr, 720, 1280, 0.075, nv12=True)
frames = torch.randint(16, 236, (12, 1080 * 1280), dtype=torch.uint8, device="cuda")
ref = pipe.run(frames, slice(2, 10)).clone()
t0 = time.time()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for i in range(reps):
    out = pipe.run(frames, slice(2, 10))
    if i % 10 == 9:
        assert torch.equal(out, ref), f"repetition {i} differs"
torch.cuda.synchronize()
print(f"cfg3 pipeline: {reps} x 8 frames, all checked repetitions identical, {reps * 8 / (time.time() - t0):.1f} frames/s incl. checks")
s = den.stream(720, 1280, in_fmt=L.FMT_NV12, noise=0.075)
first = {}
n = 0
for i in range(600):
    o = s.push(frames[i % 12].reshape(1080, 1280))
    if o is not None:
        n += 1
torch.cuda.synchronize()
rest = list(s.flush())
s.close()
print(f"stream: 600 pushes -> {n} + {len(rest)} frames")
assert n + len(rest) == 600
print("soak ok")
