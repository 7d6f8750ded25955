"""Per-helper device time of the service glue (csrc/glue.cu) for the reference's live default (720p -> x4 -> 1440p, batch 4)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ss4k_b200 import service, _lib as L
from oracle import srvgg
torch.manual_seed(0)
net = srvgg.SRVGGNetCompact(3, 3, 64, 32, 4).eval()
svc = service.FsrcnnUpscalerService(lr_level=3, device=0, denoising=False, model_name='realesr-general-x4v3', state_dict=net.state_dict(), batch_size=4)
svc.proc_init(); svc.output_shape = (1440, 2560)
frames = torch.randint(0, 256, (4, 720, 1280, 3), dtype=torch.uint8, device="cuda")
svc.upscale(frames); torch.cuda.synchronize()
times = {}
def wrap(name):
    fn = getattr(svc, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); torch.cuda.synchronize()
        times.setdefault(name, []).append(e0.elapsed_time(e1)); return r
    setattr(svc, name, w)
for n in ("_stats", "_area", "_finish", "_run_model"):
    wrap(n)
for _ in range(3):
    times.clear(); svc.upscale(frames)
print(json.dumps({k: [round(x, 3) for x in v] for k, v in times.items()}))
