"""SRVGGNetCompact restated (oracle; test infrastructure only).

Follows /root/reference/src/upscale/model/realesrgan/factory.py:18-82 :
  body[0] = Conv2d(3, nf, 3, 1, 1); body[1] = PReLU(nf); then num_conv x [Conv2d(nf, nf), PReLU(nf)];
  body[-1] = Conv2d(nf, 3*s*s); PixelShuffle(s); out += nearest_upsample(x, s).
State-dict keys are identical to the reference module's (``body.{i}.weight`` ...).
"""
import torch
from torch import nn
from torch.nn import functional as F


class SRVGGNetCompact(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type="prelu"):
        super().__init__()
        assert act_type == "prelu", "the reference only instantiates the PReLU variant (factory.py:128,132)"
        self.upscale = upscale
        layers = [nn.Conv2d(num_in_ch, num_feat, 3, 1, 1), nn.PReLU(num_parameters=num_feat)]
        for _ in range(num_conv):
            layers += [nn.Conv2d(num_feat, num_feat, 3, 1, 1), nn.PReLU(num_parameters=num_feat)]
        layers.append(nn.Conv2d(num_feat, num_out_ch * upscale * upscale, 3, 1, 1))
        self.body = nn.ModuleList(layers)

    def forward(self, x):
        out = x
        for m in self.body:
            out = m(out)
        out = F.pixel_shuffle(out, self.upscale)                                     # factory.py:78
        return out + F.interpolate(x, scale_factor=float(self.upscale), mode="nearest")  # :80-81
