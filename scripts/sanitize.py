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
# uint8 frames in / out (packed RGB row stores, ragged width), and the service glue with the SRVGG half-NCHW PixelShuffle store
net = rrdbnet.RRDBNet(3, 3, 2, 64, 1, 32).eval()
m = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=1, device=0, use_graph=False)
fr = torch.randint(0, 256, (1, 24, 150, 3), dtype=torch.uint8).cuda()
print("rrdb u8", tuple(m._plan(1, 24, 150, L.FMT_U8_NHWC, L.FMT_U8_NHWC).run(fr).shape))
from ss4k_b200 import service
sv = service.FsrcnnUpscalerService(lr_level=0, device=0, denoising=False, model_name='realesr-animevideov3',
                                   state_dict=srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval().state_dict())
sv.proc_init()
sv.output_shape = (130, 270)
print("service", tuple(sv.upscale(torch.randint(0, 256, (1, 72, 136, 3), dtype=torch.uint8).cuda()).shape))
sv.output_shape = None
print("service", tuple(sv.upscale(torch.randint(0, 256, (1, 72, 136, 3), dtype=torch.uint8).cuda()).shape))
print("rgb->nv12", tuple(ss4k_b200.Engine.get(0).rgb_to_nv12(torch.randint(0, 256, (1, 8, 32, 3), dtype=torch.uint8).cuda()).shape))
torch.cuda.synchronize()
print("done")
