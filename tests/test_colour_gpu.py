"""Colour stage at the frame boundary (north star part 4): NV12 in (fused into the layout kernel that feeds the
first conv) and RGB -> NV12 out.  The reference has no implementation of either (frames are rgb24 on its pipes), so
the oracle is this repository's own definition (oracle/colour.py): unpinned by construction."""
import math

import numpy as np
import pytest
import torch

from ss4k_b200 import _lib as L
from ss4k_b200 import realesrgan
from oracle import colour, rrdbnet

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,w", [(1, 2, 4), (2, 36, 68), (1, 720, 1280), (1, 1440, 2560)])
def test_rgb_to_nv12_bit_exact(engine, n, h, w):
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, generator=g)
    frames[0, 0, :4] = torch.tensor([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 0, 255]], dtype=torch.uint8)
    want = colour.rgb_to_nv12(frames.numpy())
    got = engine.rgb_to_nv12(frames.cuda()).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_nv12_roundtrip_property(engine):
    """Full-size property: NV12 -> (layout kernel) -> identity is not available, but RGB -> NV12 -> RGB through the
    oracle's decoder must stay within the 4:2:0 quantisation error for a smooth image."""
    h, w = 720, 1280
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    img = torch.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], dim=-1).to(torch.uint8)[None]
    nv = engine.rgb_to_nv12(img.cuda()).cpu().numpy()
    back = colour.nv12_to_rgb(nv, h, w)[0].transpose(1, 2, 0) * 255.0
    assert np.abs(back - img[0].numpy()).max() <= 3.0


def test_nv12_input_path_matches_oracle(engine):
    """NV12 frames in -> RRDBNet x2 -> uint8 RGB out; the oracle decodes NV12 with oracle/colour.py and runs the
    fp32 net."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    h, w = 40, 64
    g = torch.Generator().manual_seed(9)
    rgb = torch.randint(0, 256, (2, h, w, 3), dtype=torch.uint8, generator=g)
    nv = colour.rgb_to_nv12(rgb.numpy())
    x = torch.from_numpy(colour.nv12_to_rgb(nv, h, w))
    with torch.no_grad():
        want = net(x).clamp(0, 1)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=2, device=0)
    plan = model._plan(2, h, w, L.FMT_NV12, L.FMT_F32_NCHW)
    got = plan.run(torch.from_numpy(nv).cuda()).cpu().clamp(0, 1)
    mse = torch.mean((got - want) ** 2).item()
    psnr = 99.0 if mse == 0 else -10 * math.log10(mse)
    maxabs = (got - want).abs().max().item() * 255
    print(f"NV12 -> RRDBNet-2 x2: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0


def test_bsvd_takes_nv12_and_u8_frames(engine):
    """BSVD plans with a frame format: the layout kernel decodes NV12 (or /255 for uint8 RGB) and fills the constant
    noise-map channel (0.1 * denoise_rate, fsrcnn_upscaler.py:262); oracle = oracle/colour.py decode + oracle BSVD."""
    from ss4k_b200 import bsvd as nb
    from oracle import bsvd as ob
    sd = ob.build_bsvd32(0, weight_scale=0.5)
    t, h, w = 3, 32, 136
    g = torch.Generator().manual_seed(4)
    rgb = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, generator=g)
    nv = colour.rgb_to_nv12(rgb.numpy())
    den = nb.NativeBSVD(sd, device=0)
    for name, frames, x3 in (("nv12", torch.from_numpy(nv), torch.from_numpy(colour.nv12_to_rgb(nv, h, w))),
                             ("u8", rgb, rgb.permute(0, 3, 1, 2).float() / 255.0)):
        x = torch.cat([x3, torch.full((t, 1, h, w), 0.075)], dim=1)[None]
        want = ob.bsvd_forward(sd, x)[0].clamp(0, 1)
        got = den.denoise_frames(frames.cuda(), h, w, 0.075, nv12=(name == "nv12")).float().cpu().clamp(0, 1)
        mse = torch.mean((got - want) ** 2).item()
        psnr = 99.0 if mse == 0 else -10 * math.log10(mse)
        maxabs = (got - want).abs().max().item() * 255
        print(f"BSVD from {name} frames: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
        assert psnr >= 50 and maxabs <= 2.0


def test_nv12_surfaces_in_and_out(engine):
    """Ingest / egress surfaces: pitched per-frame NV12 surfaces as a hardware decoder hands them out
    (luma pitch 1536 for 1280 columns, chroma plane after 736 coded rows) are packed into the chunk the plans read, and
    packed NV12 frames are written into pitched encoder surfaces: byte-exact both ways, padding bytes untouched; the
    denoiser run on the packed chunk equals the run on the same frames supplied packed from the start."""
    from ss4k_b200 import bsvd as native_bsvd
    from oracle import bsvd
    h, w, n, pitch, coded_h = 72, 136, 5, 256, 80
    g = torch.Generator().manual_seed(3)
    packed = torch.randint(16, 236, (n, h * 3 // 2, w), dtype=torch.uint8, generator=g).cuda()
    pool = [torch.full((coded_h * 3 // 2, pitch), 7, dtype=torch.uint8, device="cuda") for _ in range(n)]   # decoder-owned surfaces
    surf = [(s[:h, :w], s[coded_h:coded_h + h // 2, :w]) for s in pool]
    engine.nv12_unpack(packed, surf, h, w)                     # packed -> pitched (the encoder direction)
    torch.cuda.synchronize()
    for i, s in enumerate(pool):
        assert torch.equal(s[:h, :w], packed[i, :h]) and torch.equal(s[coded_h:coded_h + h // 2, :w], packed[i, h:])
        assert (s[:h, w:] == 7).all() and (s[h:coded_h] == 7).all() and (s[coded_h:, w:] == 7).all()   # padding untouched
    again = engine.nv12_pack(surf, h, w)                       # pitched -> packed (the decoder direction)
    assert torch.equal(again, packed)
    den = native_bsvd.NativeBSVD(bsvd.build_bsvd32(0, weight_scale=0.5), device=0)
    a = den.denoise_frames(again.reshape(n, -1), h, w, 0.075, nv12=True)
    b = den.denoise_frames(packed.reshape(n, -1), h, w, 0.075, nv12=True)
    assert torch.equal(a, b)
    with pytest.raises(L.Ss4kError):
        engine.nv12_pack([(pool[0][:h, :w - 8], pool[0][coded_h:coded_h + h // 2, :w])], h, w + 512)   # pitch < width
