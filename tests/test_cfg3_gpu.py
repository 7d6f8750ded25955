"""The headline configuration (BASELINE.json configs[2]) on the GPU, full 1280x720 frames:
NV12 stream -> BSVD-32 temporal denoiser over the chunk (reference constructor init, fp16 hi/lo split precision)
-> sharpen / clamp / 0.8-0.2 blend with the decoded frame (fsrcnn_upscaler.py:278-281) -> RRDBNet-23 x2 on the owned frame
-> 2560x1440, through ``pipeline.DenoiseUpscalePipeline`` (the call bench.py times),
against the CPU oracle chain oracle.colour -> oracle.bsvd -> oracle.rrdbnet on the same seeded frames and weights.

Gate (BASELINE.json north_star): PSNR >= 50 dB and max |err| <= 2/255 on clamped [0,1] RGB, compared BEFORE the uint8
quantisation (half NCHW output); the uint8 frame must equal the quantised float result of the same engine run within
1 LSB (truncation of a value that sits on an integer boundary)."""
import math
import os

import numpy as np
import pytest
import torch

from ss4k_b200 import _lib as L
from ss4k_b200 import bsvd as native_bsvd
from ss4k_b200 import realesrgan, sharding
from ss4k_b200.pipeline import DenoiseUpscalePipeline
from oracle import bsvd, colour, glue, rrdbnet

pytestmark = pytest.mark.gpu

H, W = 720, 1280
NOISE = 0.075


def _nv12_stream(t, seed=1234):
    """smooth moving pattern + gaussian noise, encoded with the oracle's RGB -> NV12"""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    frames = np.empty((t, H * W * 3 // 2), dtype=np.uint8)
    for i in range(t):
        rgb = np.stack([(xx + 4 * i) % 256, (yy + 2 * i) % 256, (xx + yy) // 8 % 256], axis=-1).astype(np.float32)
        rgb = np.clip(rgb + g.normal(0, 10, rgb.shape), 0, 255).astype(np.uint8)
        frames[i] = colour.rgb_to_nv12(rgb[None])[0]
    return frames


def _gate(got, want):
    a, b = got.float().cpu().clamp(0, 1), want.clamp(0, 1)
    mse = torch.mean((a - b) ** 2).item()
    return (99.0 if mse == 0 else -10 * math.log10(mse)), (a - b).abs().max().item() * 255


@pytest.fixture(scope="module")
def nets(engine):
    torch.manual_seed(0)
    rr = rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval()
    bsd = bsvd.build_bsvd32(0)                      # the reference constructor's init -> split precision
    sr = realesrgan.NativeRRDBNet(rr.state_dict(), scale=2, num_block=23, device=0)
    den = native_bsvd.NativeBSVD(bsd, device=0, act_mode="auto", out_dtype=torch.float16)
    assert den.act_mode == L.ACT_F16_SPLIT
    return rr, bsd, sr, den


def test_cfg3_720p_chunk_vs_oracle(nets):
    rr, bsd, sr, den = nets
    T, own = 3, slice(1, 2)
    nv = _nv12_stream(T)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        rgb = torch.from_numpy(colour.nv12_to_rgb(nv, H, W))                        # [T,3,H,W]
        x = torch.cat([rgb, torch.full((T, 1, H, W), NOISE)], dim=1)[None]          # [1,T,4,H,W]
        den_want = bsvd.bsvd_forward(bsd, x)[0]                                      # [T,3,H,W]
        # glue between the nets (fsrcnn_upscaler.py:278-281): sharpen(2e-5) + clamp, 0.8 * den + 0.2 * original
        d = den_want[own]
        d = torch.clamp(glue.depthwise_reflect(d.reshape(-1, 1, H, W), glue.sharpen_weight(0.00002)).reshape(d.shape), 0, 1)
        want = rr(d * 0.8 + 0.2 * rgb[own])
    frames = torch.from_numpy(nv).cuda()
    pf = DenoiseUpscalePipeline(den, sr, H, W, NOISE, nv12=True, out_fmt=L.FMT_F16_NCHW)
    got = pf.run(frames, own)
    torch.cuda.synchronize()
    assert tuple(got.shape) == (1, 3, 2 * H, 2 * W)
    psnr, maxabs = _gate(got, want)
    print(f"cfg3 720p (NV12 -> BSVD split F={T} -> RRDBNet-23 x2), before quantisation: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0
    # the uint8 path of the same pipeline (what bench.py times)
    pu = DenoiseUpscalePipeline(den, sr, H, W, NOISE, nv12=True, out_fmt=L.FMT_U8_NHWC)
    u8 = pu.run(frames, own)
    torch.cuda.synchronize()
    q = (got.float().clamp(0, 1) * 255).permute(0, 2, 3, 1)
    d = (u8.float() - q.floor()).abs()
    assert d.max().item() <= 1.0
    want_u8 = (want.clamp(0, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1)
    du = (u8.cpu().int() - want_u8.int()).abs()
    print(f"  uint8 frame vs truncated oracle: max {int(du.max())} LSB, mean {du.float().mean().item():.4f} LSB")
    assert du.max().item() <= 3 and du.float().mean().item() <= 0.6


def test_cfg3_chunk_halo_equals_whole_stream(nets):
    """sharding.bsvd_chunks: chunk + 16-frame halo per rank reproduces the frames of the un-sharded stream bit for bit
    (small frames: the property is size independent; world 2 and 3)."""
    rr, bsd, sr, den = nets
    h, w, T = 64, 128, 40
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (T, h * w * 3 // 2), dtype=torch.uint8, generator=g).cuda()
    pipe = DenoiseUpscalePipeline(den, sr, h, w, NOISE, nv12=True, out_fmt=L.FMT_U8_NHWC)
    whole = pipe.run(frames).clone()
    for world in (2, 3):
        parts = []
        for ch in sharding.bsvd_chunks(T, world):
            parts.append(pipe.run(frames[ch.load_lo:ch.load_hi], ch.owned).clone())
        assert torch.equal(torch.cat(parts, dim=0), whole)


def test_cfg3_host_path_equals_device_path(nets):
    rr, bsd, sr, den = nets
    h, w, T = 64, 128, 6
    g = torch.Generator().manual_seed(9)
    pipe = DenoiseUpscalePipeline(den, sr, h, w, NOISE, nv12=True, out_fmt=L.FMT_U8_NHWC)
    clips = [torch.randint(0, 256, (T, h * w * 3 // 2), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    own = slice(1, 5)
    want = [pipe.run(c.cuda(), own).cpu() for c in clips]
    outs = [torch.empty((4,) + pipe.out_frame_shape(), dtype=torch.uint8).pin_memory() for _ in range(3)]
    for c, o in zip(clips, outs):
        pipe.run_host(c, own, o)
    pipe.host_sync()
    for a, b in zip(want, outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("nv12", [True, False])
def test_cfg3_glue_writes_the_upscalers_first_tensor(nets, nv12):
    """The hand-over between the nets: the sharpen / clamp / blend glue writes RRDBNet's pixel-unshuffled fp16 activation
    tensor itself (ss4k_glue_sharpen_blend_act + ss4k_run_act: no float image, no layout kernel in the upscaler's plan).
    Must equal the two-step path (float NCHW image -> the plan's own layout kernel) bit for bit."""
    rr, bsd, sr, den = nets
    h, w, T = 72, 136, 5
    g = torch.Generator().manual_seed(13)
    shape = (T, h * w * 3 // 2) if nv12 else (T, h, w, 3)
    frames = torch.randint(16, 236, shape, dtype=torch.uint8, generator=g).cuda()
    pipe = DenoiseUpscalePipeline(den, sr, h, w, NOISE, nv12=nv12, out_fmt=L.FMT_U8_NHWC)
    assert pipe._act is not None
    direct = pipe.run(frames, slice(1, 4)).clone()
    pipe._act = None
    two_step = pipe.run(frames, slice(1, 4))
    assert torch.equal(direct, two_step)
