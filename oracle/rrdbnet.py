"""basicsr RRDBNet restated (oracle; test infrastructure only).

The reference imports it from the un-vendored pip package ``basicsr``
(/root/reference/src/upscale/model/realesrgan/factory.py:6, constructors at :113,117,121,125;
forward reached at src/upscale/fsrcnn_upscaler.py:181,294).  Restated from the published
``basicsr/archs/rrdbnet_arch.py`` + ``arch_util.py``:
  RDB : x1..x4 = lrelu(conv_k(cat(x, x1..x_{k-1}))), x5 = conv5(cat(...)); return x5*0.2 + x
  RRDB: rdb3(rdb2(rdb1(x)))*0.2 + x
  net : [pixel_unshuffle(2) if scale==2] -> conv_first -> body -> conv_body (+skip) ->
        2x (nearest x2 -> conv_up -> lrelu) -> conv_hr -> lrelu -> conv_last
  init: default_init_weights(RDB convs, scale=0.1) = kaiming_normal_ (fan_in, a=0) * 0.1, bias 0.
Known-answer pins available without the package: parameter counts 16,697,987 (x4, 23 blocks),
16,703,171 (x2), 4,467,779 (x4, 6 blocks) and the state-dict key list.  PARITY UNPINNED otherwise.
"""
import torch
from torch import nn
from torch.nn import functional as F


def pixel_unshuffle(x, scale):
    b, c, hh, hw = x.size()
    out_channel = c * (scale ** 2)
    assert hh % scale == 0 and hw % scale == 0
    h, w = hh // scale, hw // scale
    x_view = x.view(b, c, h, scale, w, scale)
    return x_view.permute(0, 1, 3, 5, 2, 4).reshape(b, out_channel, h, w)


def default_init_weights(module_list, scale=1.0):
    for module in module_list:
        for m in module.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.fill_(0)


class ResidualDenseBlock(nn.Module):
    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        default_init_weights([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5], 0.1)

    def forward(self, x):
        x1 = self.lrelu(self.conv1(x))
        x2 = self.lrelu(self.conv2(torch.cat((x, x1), 1)))
        x3 = self.lrelu(self.conv3(torch.cat((x, x1, x2), 1)))
        x4 = self.lrelu(self.conv4(torch.cat((x, x1, x2, x3), 1)))
        x5 = self.conv5(torch.cat((x, x1, x2, x3, x4), 1))
        return x5 * 0.2 + x


class RRDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)

    def forward(self, x):
        out = self.rdb3(self.rdb2(self.rdb1(x)))
        return out * 0.2 + x


class RRDBNet(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        if scale == 2:
            num_in_ch = num_in_ch * 4
        elif scale == 1:
            num_in_ch = num_in_ch * 16
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        if self.scale == 2:
            feat = pixel_unshuffle(x, 2)
        elif self.scale == 1:
            feat = pixel_unshuffle(x, 4)
        else:
            feat = x
        feat = self.conv_first(feat)
        body_feat = self.conv_body(self.body(feat))
        feat = feat + body_feat
        feat = self.lrelu(self.conv_up1(F.interpolate(feat, scale_factor=2, mode="nearest")))
        feat = self.lrelu(self.conv_up2(F.interpolate(feat, scale_factor=2, mode="nearest")))
        return self.conv_last(self.lrelu(self.conv_hr(feat)))


def tile_process(model, img, scale, tile, tile_pad):
    """RealESRGANer.tile_process semantics (realesrgan/utils.py @5ca1078):
    crop each tile expanded by tile_pad (clamped to the image), run, paste the un-padded centre."""
    import math
    b, c, h, w = img.shape
    out = img.new_zeros(b, c, h * scale, w * scale)
    tiles_x, tiles_y = math.ceil(w / tile), math.ceil(h / tile)
    for y in range(tiles_y):
        for x in range(tiles_x):
            sx, sy = x * tile, y * tile
            ex, ey = min(sx + tile, w), min(sy + tile, h)
            sxp, exp_ = max(sx - tile_pad, 0), min(ex + tile_pad, w)
            syp, eyp = max(sy - tile_pad, 0), min(ey + tile_pad, h)
            o = model(img[:, :, syp:eyp, sxp:exp_])
            ox0, oy0 = (sx - sxp) * scale, (sy - syp) * scale
            out[:, :, sy * scale:ey * scale, sx * scale:ex * scale] = \
                o[:, :, oy0:oy0 + (ey - sy) * scale, ox0:ox0 + (ex - sx) * scale]
    return out


def enhance_tensor(model, img, scale, tile=0, tile_pad=10, pre_pad=0):
    """RealESRGANer.pre_process -> process / tile_process -> post_process on a float tensor [B,3,H,W]
    (realesrgan/utils.py @5ca1078): reflect pre_pad on the right / bottom, reflect mod-2 pad for the
    x2 net, tiles with tile_pad, then both pads cropped off the output (x scale)."""
    import torch.nn.functional as F
    if pre_pad:
        img = F.pad(img, (0, pre_pad, 0, pre_pad), "reflect")
    mod_h = mod_w = 0
    if scale == 2:
        _, _, h, w = img.shape
        mod_h, mod_w = (2 - h % 2) % 2, (2 - w % 2) % 2
        if mod_h or mod_w:
            img = F.pad(img, (0, mod_w, 0, mod_h), "reflect")
    out = tile_process(model, img, scale, tile, tile_pad) if tile > 0 else model(img)
    _, _, h, w = out.shape
    out = out[:, :, 0:h - mod_h * scale, 0:w - mod_w * scale]
    _, _, h, w = out.shape
    return out[:, :, 0:h - pre_pad * scale, 0:w - pre_pad * scale]
