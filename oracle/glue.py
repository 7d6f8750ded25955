"""Service glue restated as pure functions (oracle; test infrastructure only).

Follows /root/reference/src/upscale/fsrcnn_upscaler.py line by line, keeping its quirks (PINNED: reproduces the uint8
outputs of the reference's own code bit for bit, tests/golden/glue.npz, tests/test_oracle_cpu.py::test_glue_golden):
  blur_ker :20-52, sharpen_ker :54-84, upscale_multi :168-233, upscale_single :235-326
  (unbiased std, +1e-8, bicubic-always resize because ``output_shape[0] >= batch dim``, truncating uint8).
fp16 autocast of the reference is NOT emulated: the oracle is the fp32 meaning of the same graph.
"""
import math

import torch
import torch.nn.functional as F


def blur_weight(kernel_size=3, sigma=0.5):
    x_cord = torch.arange(kernel_size)
    x_grid = x_cord.repeat(kernel_size).view(kernel_size, kernel_size)
    y_grid = x_grid.t()
    xy_grid = torch.stack([x_grid, y_grid], dim=-1)
    mean = (kernel_size - 1) / 2.0
    variance = sigma ** 2.0
    g = (1.0 / (2.0 * math.pi * variance)) * torch.exp(-torch.sum((xy_grid - mean) ** 2.0, dim=-1) / (2 * variance))
    g = g / torch.sum(g)
    return g.view(1, 1, kernel_size, kernel_size).float()


def sharpen_weight(strength=1.0):
    sharp = torch.tensor([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]], dtype=torch.float32)
    ident = torch.tensor([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=torch.float32)
    k = sharp * strength + (1 - strength) * ident
    k = k / torch.sum(k)
    return k.view(1, 1, 3, 3)


def depthwise_reflect(x, weight):
    """x [N,1,H,W]; Conv2d(groups=1 channel, padding_mode='reflect')"""
    k = weight.shape[-1]
    return F.conv2d(F.pad(x, (k // 2,) * 4, mode="reflect"), weight)


def match_mean_std(hr, lr):
    """fsrcnn_upscaler.py:188-199 -- per (N, C) mean / unbiased std match of hr to lr."""
    n, c, h, w = hr.shape
    hm = hr.reshape(n, c, -1).mean(dim=-1).view(n, c, 1, 1)
    hs = hr.reshape(n, c, -1).std(dim=-1).view(n, c, 1, 1)
    lm = lr.reshape(n, c, -1).mean(dim=-1).view(n, c, 1, 1)
    ls = lr.reshape(n, c, -1).std(dim=-1).view(n, c, 1, 1)
    return (hr - hm) / (hs + 1e-8) * ls + lm


def match_local_colour(hr, lr, factor=8):
    """fsrcnn_upscaler.py:201-218."""
    n, c, h, w = hr.shape
    mb = blur_weight(kernel_size=17, sigma=8.0)
    if (h // factor) > (mb.shape[-1] // 2) and h > 64 and w > 64:
        size = (h // factor, w // factor)
        lb = F.interpolate(lr, size=size, mode="area")
        hb = F.interpolate(hr, size=size, mode="area")
        lb = depthwise_reflect(lb.reshape(n * c, 1, *size), mb).reshape(n, c, *size)
        hb = depthwise_reflect(hb.reshape(n * c, 1, *size), mb).reshape(n, c, *size)
        diff = F.interpolate(hb - lb, size=(h, w), mode="bilinear")
        hr = hr - diff
    return hr


def upscale_multi(frames_u8, model, lr_shape, output_shape=None, lr_hr_resize=True, return_float=False):
    """frames_u8 [N,H,W,3] uint8 -> uint8 [N,H',W',3]  (fsrcnn_upscaler.py:168-233).
    return_float: also return the float frame (x 255, NHWC) in front of the truncating ``.to(torch.uint8)``."""
    with torch.no_grad():
        img = frames_u8.permute(0, 3, 1, 2) / 255.0
        lr = img
        if (img.shape[-1] > lr_shape[-1] or img.shape[-2] > lr_shape[-2]) and lr_hr_resize:
            lr = F.interpolate(img, size=lr_shape, mode="area")
        hr = model(lr).float()
        hr = match_mean_std(hr, lr)
        hr = match_local_colour(hr, lr)
        hr = torch.clamp(hr, 0, 1)
        if output_shape is not None and lr_hr_resize:
            # reference compares output_shape[0] with the BATCH dim (:223) -> bicubic in practice
            mode = "bicubic" if output_shape[0] >= hr.shape[0] else "area"
            hr = F.interpolate(hr, size=output_shape, mode=mode)
        hr = torch.clamp(hr, 0, 1)
        q = (hr * 255).permute(0, 2, 3, 1)
        return (q.to(torch.uint8), q) if return_float else q.to(torch.uint8)


def denoise_frame(frame_chw, denoise_model, denoise_rate, first_frame):
    """Denoise branch of upscale_single (fsrcnn_upscaler.py:245-284) for ONE frame (F=1 clip).
    frame_chw: float [3,H,W] in [0,1]; denoise_model: callable [1,1,4,H,W] -> [1,1,3,H,W]."""
    c, h, w = frame_chw.shape
    x = torch.empty(1, 1, 4, h, w)
    x[0, 0, :3] = frame_chw
    x[0, 0, 3] = 0.05 if first_frame else 0.1 * denoise_rate       # :262,269
    den = denoise_model(x)[:, -1][0]
    den = torch.clamp(depthwise_reflect(den.view(3, 1, h, w), sharpen_weight(0.00002)).view(3, h, w), 0, 1)
    return den * 0.8 + 0.2 * frame_chw                              # :281


def upscale_single(frame_u8, model, lr_shape, output_shape=None, denoise_model=None, denoise_rate=1.0,
                   first_frame=True, return_float=False):
    """frame_u8 [H,W,3] uint8 -> uint8 [H',W',3]  (fsrcnn_upscaler.py:235-326, realesrgan branch).
    denoise_model: callable [1,1,4,H,W] -> [1,1,3,H,W] or None (denoising=False)."""
    with torch.no_grad():
        img = frame_u8.permute(2, 0, 1).unsqueeze(0) / 255.0
        lr_before = F.interpolate(img, size=lr_shape, mode="area").squeeze(0)            # :237-241
        lr = lr_before
        if denoise_model is not None:
            lr = denoise_frame(lr_before, denoise_model, denoise_rate, first_frame)      # :245-284
        hr = model(lr.unsqueeze(0).float())[0]                                           # :293-295
        c, h, w = hr.shape
        if denoise_model is not None:
            hr = torch.clamp(depthwise_reflect(hr.view(c, 1, h, w), sharpen_weight(0.00007)).view(c, h, w), 0, 1)  # :298-299
        hm = hr.reshape(c, -1).mean(dim=-1).view(c, 1, 1)                                 # :302-313
        hs = hr.reshape(c, -1).std(dim=-1).view(c, 1, 1)
        lm = lr_before.reshape(c, -1).mean(dim=-1).view(c, 1, 1)
        ls = lr_before.reshape(c, -1).std(dim=-1).view(c, 1, 1)
        hr = (hr - hm) / (hs + 1e-8) * ls + lm
        out = torch.clamp(hr, 0, 1).unsqueeze(0)
        if output_shape is not None:
            out = F.interpolate(out, size=output_shape, mode="bicubic")                  # :316-325 (bicubic in practice)
        out = torch.clamp(out, 0, 1)
        q = (out * 255)[0].permute(1, 2, 0)
        return (q.to(torch.uint8), q) if return_float else q.to(torch.uint8)
