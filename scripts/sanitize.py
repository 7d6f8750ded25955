"""Small end-to-end runs for compute-sanitizer (memcheck): SRVGG, RRDBNet (dense slab, residual epilogues, TMA stores),
BSVD clip (stride-2 tile kernel, scatter stores, split mode)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L, realesrgan, bsvd as nb
from oracle import rrdbnet, srvgg, bsvd as ob
torch.manual_seed(0)
x = torch.rand(1, 3, 24, 136).cuda()
m = realesrgan.NativeSRVGG(srvgg.SRVGGNetCompact(3, 3, 64, 2, 4).eval().state_dict(), num_conv=2, upscale=4, device=0, use_graph=False)
print("srvgg", tuple(m(x).shape))
m = realesrgan.NativeRRDBNet(rrdbnet.RRDBNet(3, 3, 2, 64, 1, 32).eval().state_dict(), scale=2, num_block=1, device=0, use_graph=False)
print("rrdb", tuple(m(x).shape))
for mode in (L.ACT_F16, L.ACT_F16_SPLIT):
    d = nb.NativeBSVD(ob.build_bsvd32(0, weight_scale=0.5), device=0, act_mode=mode, use_graph=False)
    print("bsvd", mode, tuple(d(torch.rand(1, 2, 4, 16, 136).cuda()).shape))
print("rgb->nv12", tuple(ss4k_b200.Engine.get(0).rgb_to_nv12(torch.randint(0, 256, (1, 8, 32, 3), dtype=torch.uint8).cuda()).shape))
torch.cuda.synchronize()
print("done")
