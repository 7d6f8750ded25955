"""One SRVGGNetCompact-32 x4 frame (1280x720) on the engine: target of the ncu capture of the 64-wide and 48-wide variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L, realesrgan
from oracle import srvgg
torch.manual_seed(0)
net = srvgg.SRVGGNetCompact(3, 3, 64, 32, 4).eval()
m = realesrgan.NativeSRVGG(net.state_dict(), num_conv=32, upscale=4, device=0, out_dtype=torch.float16)
x = torch.rand(1, 3, 720, 1280, device="cuda")
for _ in range(2):
    y = m(x)
torch.cuda.synchronize()
print(tuple(y.shape))
