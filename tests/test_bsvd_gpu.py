"""GPU parity of the BSVD denoiser through the drop-in model interface against the CPU oracle clip function
(pinned to the reference's BSVD.forward by tests/test_oracle_cpu.py).
Gate (BASELINE.json north_star): PSNR >= 50 dB and max |err| <= 2/255 on clamped [0,1] RGB."""
import math

import pytest
import torch

from ss4k_b200 import _lib as L
from ss4k_b200 import bsvd as native_bsvd
from oracle import bsvd

pytestmark = pytest.mark.gpu


def gate(got, want):
    a, b = got.float().cpu().clamp(0, 1), want.clamp(0, 1)
    mse = torch.mean((a - b) ** 2).item()
    psnr = 99.0 if mse == 0 else -10 * math.log10(mse)
    return psnr, (a - b).abs().max().item() * 255


def _clip(frames, h, w, seed=1234):
    x = torch.rand(1, frames, 4, h, w, generator=torch.Generator().manual_seed(seed))
    x[:, :, 3] = 0.075   # noise map = 0.1 * denoise_rate (fsrcnn_upscaler.py:262), denoise_rate 0.75
    return x


@pytest.mark.parametrize("frames,h,w", [(1, 72, 128), (5, 64, 136), (3, 36, 260)])
def test_bsvd_trained_like_weights(engine, frames, h, w):
    """Constructor init scaled by 0.5 ('trained-like' magnitudes): single-MMA fp16 operands."""
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    x = _clip(frames, h, w)
    want = bsvd.bsvd_forward(sd, x)
    model = native_bsvd.NativeBSVD(sd, device=0)
    got = model(x.cuda())
    torch.cuda.synchronize()
    assert tuple(got.shape) == (1, frames, 3, h, w)
    psnr, maxabs = gate(got, want)
    print(f"BSVD-32 (weights x0.5) F={frames} {h}x{w}: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0


def test_bsvd_clip_is_temporal(engine):
    """A frame inside a clip must differ from the same frame denoised alone, and repeated
    calls must be identical (state reset, model.py:579)."""
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    x = _clip(4, 32, 128, seed=7).cuda()
    model = native_bsvd.NativeBSVD(sd, device=0)
    full = model(x).clone()
    again = model(x).clone()
    single = model(x[:, 1:2]).clone()
    assert torch.equal(full, again)
    assert (full[:, 1] - single[:, 0]).abs().max().item() > 1e-3


def test_bsvd_constructor_init_report(engine):
    """The reference constructor's kaiming init gives outputs spanning +-14: single-MMA fp16
    is expected to miss max-abs there; report the numbers, gate only PSNR >= 45 dB."""
    sd = bsvd.build_bsvd32(0)
    x = _clip(3, 48, 128)
    want = bsvd.bsvd_forward(sd, x)
    got = native_bsvd.NativeBSVD(sd, device=0)(x.cuda())
    psnr, maxabs = gate(got, want)
    rel = (got.float().cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"BSVD-32 (constructor init) fp16: PSNR {psnr:.1f} dB, max|err| {maxabs:.2f}/255, rel {rel:.2e}")
    assert psnr >= 45 and rel < 5e-3


@pytest.mark.parametrize("frames,h,w", [(3, 48, 128), (1, 72, 128), (5, 64, 136)])
def test_bsvd_constructor_init_split_precision(engine, frames, h, w):
    """Reference constructor init (kaiming_normal_, model.py:393-400,501-508; fp32 outputs span +-14) through the
    fp16 hi/lo split mode (3 MMAs per product, every activation tensor stored as hi + lo): full north-star gate."""
    sd = bsvd.build_bsvd32(0)
    x = _clip(frames, h, w)
    want = bsvd.bsvd_forward(sd, x)
    model = native_bsvd.NativeBSVD(sd, device=0, act_mode=L.ACT_F16_SPLIT)
    got = model(x.cuda())
    torch.cuda.synchronize()
    psnr, maxabs = gate(got, want)
    rel = (got.float().cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"BSVD-32 (constructor init) fp16 split F={frames} {h}x{w}: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255, rel {rel:.2e}")
    assert psnr >= 50 and maxabs <= 2.0


def test_bsvd_auto_precision(engine):
    """act_mode='auto' (the build_model default) picks the split mode for kaiming-magnitude weights and the
    single-MMA mode for trained-like magnitudes."""
    assert native_bsvd.NativeBSVD(bsvd.build_bsvd32(0), device=0, act_mode="auto").act_mode == L.ACT_F16_SPLIT
    assert native_bsvd.NativeBSVD(bsvd.build_bsvd32(0, weight_scale=0.5), device=0, act_mode="auto").act_mode == L.ACT_F16


def test_bsvd_streaming_equals_clip(engine):
    """Ring-buffer streaming (one frame per push, 16 frames of latency, flush at the end) must reproduce the clip
    result; a second clip after reset() must not see state of the first (model.py:579)."""
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    model = native_bsvd.NativeBSVD(sd, device=0)
    x = _clip(21, 32, 136, seed=11).cuda()
    clip = model(x)[0]
    s = model.stream(32, 136)
    assert s.latency == 16
    outs = []
    for i in range(21):
        o = s.push(x[0, i])
        assert (o is None) == (i < 16)
        if o is not None:
            outs.append(o)
    outs += list(s.flush())
    got = torch.cat(outs, dim=0)
    assert tuple(got.shape) == (21, 3, 32, 136)
    assert (got - clip).abs().max().item() <= 1e-3
    s.reset()
    x2 = _clip(3, 32, 136, seed=12).cuda()
    got2 = torch.cat([o for o in (s.push(x2[0, i]) for i in range(3)) if o is not None] + list(s.flush()), dim=0)
    assert (got2 - model(x2)[0]).abs().max().item() <= 1e-3
    assert (model.streaming_forward(x2[0]) - got2).abs().max().item() == 0.0


def test_bsvd_streaming_split_precision(engine):
    """The ring-buffer engine in the split precision mode (every ring has a low-half twin): reference constructor init
    (fp32 outputs span +-14), frame-at-a-time == the clip program bit for bit, and the north-star gate against the
    oracle's BSVD.forward over the whole clip (model.py:515-580: streaming_forward == forward)."""
    sd = bsvd.build_bsvd32(0)
    model = native_bsvd.NativeBSVD(sd, device=0, act_mode=L.ACT_F16_SPLIT)
    x = _clip(19, 32, 136, seed=21)
    clip = model(x.cuda())[0]
    s = model.stream(32, 136)
    outs = []
    for i in range(19):
        o = s.push(x[0, i].cuda())
        assert (o is None) == (i < 16)
        if o is not None:
            outs.append(o)
    outs += list(s.flush())
    got = torch.cat(outs, dim=0)
    assert tuple(got.shape) == (19, 3, 32, 136)
    assert torch.equal(got, clip)
    psnr, maxabs = gate(got, bsvd.bsvd_forward(sd, x)[0])
    print(f"BSVD-32 streaming, split precision, F=19: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0
    s.reset()   # a second, shorter clip (T < 16: the drain alone produces every frame)
    x2 = _clip(3, 32, 136, seed=22).cuda()
    got2 = torch.cat([o for o in (s.push(x2[0, i]) for i in range(3)) if o is not None] + list(s.flush()), dim=0)
    assert torch.equal(got2, model(x2)[0])
    s.close()


@pytest.mark.parametrize("nv12", [False, True])
def test_bsvd_streaming_frame_formats(engine, nv12):
    """uint8 RGB / NV12 frames pushed one at a time (the layout kernel decodes, normalises and appends the noise map)
    == the frame-format clip entry denoise_frames() on the same frames."""
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    model = native_bsvd.NativeBSVD(sd, device=0)
    h, w, t = 32, 128, 18
    g = torch.Generator().manual_seed(5)
    if nv12:
        frames = torch.randint(16, 236, (t, h * 3 // 2, w), generator=g, dtype=torch.uint8).cuda()
    else:
        frames = torch.randint(0, 256, (t, h, w, 3), generator=g, dtype=torch.uint8).cuda()
    clip = model.denoise_frames(frames, h, w, 0.075, nv12=nv12)
    s = model.stream(h, w, in_fmt=L.FMT_NV12 if nv12 else L.FMT_U8_NHWC, noise=0.075)
    outs = [o for o in (s.push(frames[i]) for i in range(t)) if o is not None] + list(s.flush())
    s.close()
    got = torch.cat(outs, dim=0)
    assert tuple(got.shape) == tuple(clip.shape) == (t, 3, h, w)
    assert (got - clip.float()).abs().max().item() <= 1e-3


@pytest.mark.parametrize("nv12", [False, True])
@pytest.mark.parametrize("split", [False, True])
def test_bsvd_first_layer_decodes_frames(engine, nv12, split, monkeypatch):
    """North-star part 4: with uint8 RGB / NV12 frames the clip plan has no layout kernel -- the first conv's producer
    warp decodes the frames (BT.709 / 255, noise-map channel) into its activation slabs and leaves the 16-bit rows for the
    DenBlock's residual.  Must equal the plan with the stand-alone layout kernel (SS4K_NO_FUSED_PREP=1) bit for bit:
    ragged width (three strips, the last one partial), several row bands per CTA, owned-range plans with a temporal halo."""
    sd = bsvd.build_bsvd32(0) if split else bsvd.build_bsvd32(0, weight_scale=0.5)
    mode = L.ACT_F16_SPLIT if split else L.ACT_F16
    t, h, w = 6, 36, 264
    g = torch.Generator().manual_seed(31)
    if nv12:
        frames = torch.randint(16, 236, (t, h * 3 // 2, w), generator=g, dtype=torch.uint8).cuda()
    else:
        frames = torch.randint(0, 256, (t, h, w, 3), generator=g, dtype=torch.uint8).cuda()
    fused = native_bsvd.NativeBSVD(sd, device=0, act_mode=mode)
    a = fused.denoise_frames(frames, h, w, 0.075, nv12=nv12).clone()
    a_own = fused.denoise_frames(frames, h, w, 0.075, nv12=nv12, own=(2, 5)).clone()
    monkeypatch.setenv("SS4K_NO_FUSED_PREP", "1")
    plain = native_bsvd.NativeBSVD(sd, device=0, act_mode=mode)
    b = plain.denoise_frames(frames, h, w, 0.075, nv12=nv12)
    b_own = plain.denoise_frames(frames, h, w, 0.075, nv12=nv12, own=(2, 5))
    torch.cuda.synchronize()
    assert fused._plan(t, h, w, L.FMT_NV12 if nv12 else L.FMT_U8_NHWC, L.FMT_F32_NCHW, 0.075).launches == \
        plain._plan(t, h, w, L.FMT_NV12 if nv12 else L.FMT_U8_NHWC, L.FMT_F32_NCHW, 0.075).launches - 1
    assert torch.equal(a, b)
    assert torch.equal(a_own, b_own) and torch.equal(a_own, a[2:5])


@pytest.mark.parametrize("split", [False, True])
def test_bsvd_streaming_steady_state_graphs(engine, split, monkeypatch):
    """A steady-state push (all layers active, no clip boundary in reach) is ONE graph launch between two copies: ring
    lengths are powers of two, so the launch parameters depend only on t mod period and one graph per phase is captured
    the first time the phase comes up.  60 frames: every phase is captured and re-used; equals the clip program and the
    un-graphed stream (SS4K_NO_STREAM_GRAPH=1) bit for bit."""
    sd = bsvd.build_bsvd32(0) if split else bsvd.build_bsvd32(0, weight_scale=0.5)
    model = native_bsvd.NativeBSVD(sd, device=0, act_mode=L.ACT_F16_SPLIT if split else L.ACT_F16)
    x = _clip(60, 32, 136, seed=41).cuda()
    clip = model(x)[0]

    def run_stream():
        s = model.stream(32, 136)
        before = engine.launch_count
        outs = [o for o in (s.push(x[0, i]) for i in range(60)) if o is not None] + list(s.flush())
        s.close()
        return torch.cat(outs, dim=0), engine.launch_count - before

    graphed, _ = run_stream()
    monkeypatch.setenv("SS4K_NO_STREAM_GRAPH", "1")
    plain, _ = run_stream()
    assert torch.equal(graphed, plain)
    assert (graphed - clip).abs().max().item() <= 1e-3
    if split:
        assert torch.equal(graphed, clip)
