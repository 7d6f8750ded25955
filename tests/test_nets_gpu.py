"""GPU parity of whole networks through the drop-in model interface against the CPU oracle.
Gate (BASELINE.json north_star): PSNR >= 50 dB and max |err| <= 2/255 on clamped [0,1] RGB."""
import math

import pytest
import torch

from ss4k_b200 import _lib as L
from ss4k_b200 import realesrgan
from oracle import rrdbnet, srvgg

pytestmark = pytest.mark.gpu


def gate(got, want):
    a, b = got.float().cpu().clamp(0, 1), want.clamp(0, 1)
    mse = torch.mean((a - b) ** 2).item()
    psnr = 99.0 if mse == 0 else -10 * math.log10(mse)
    maxabs = (a - b).abs().max().item() * 255
    return psnr, maxabs


def test_srvgg_cfg1(engine):
    """BASELINE.json configs[0]: SRVGGNetCompact-32 x4 on one 320x180 RGB frame, random init (PyTorch default)."""
    torch.manual_seed(0)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 32, 4).eval()
    x = torch.rand(1, 3, 180, 320, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        want = net(x)
    model = realesrgan.NativeSRVGG(net.state_dict(), num_conv=32, upscale=4, device=0)
    got = model(x.cuda())
    torch.cuda.synchronize()
    assert tuple(got.shape) == (1, 3, 720, 1280)
    psnr, maxabs = gate(got, want)
    print(f"SRVGG-32 x4 180x320: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0


def test_srvgg_batch_and_ragged(engine):
    torch.manual_seed(1)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval()
    x = torch.rand(3, 3, 37, 131)   # odd sizes, width just over one tile
    with torch.no_grad():
        want = net(x)
    model = realesrgan.NativeSRVGG(net.state_dict(), num_conv=16, upscale=4, device=0)
    got = model(x.cuda())
    psnr, maxabs = gate(got, want)
    assert psnr >= 50 and maxabs <= 2.0
    # graph replay: a second call with other data must not reuse stale results
    x2 = torch.rand(3, 3, 37, 131)
    with torch.no_grad():
        want2 = net(x2)
    psnr2, maxabs2 = gate(model(x2.cuda()), want2)
    assert psnr2 >= 50 and maxabs2 <= 2.0


@pytest.mark.parametrize("scale,blocks,h,w", [(2, 23, 96, 128), (4, 6, 48, 64), (2, 2, 50, 262)])
def test_rrdbnet(engine, scale, blocks, h, w):
    """RRDBNet with the upstream init (RDB convs kaiming*0.1, bias 0), fp16 operands."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, scale, 64, blocks, 32).eval()
    x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        want = net(x)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=scale, num_block=blocks, device=0)
    got = model(x.cuda())
    torch.cuda.synchronize()
    assert tuple(got.shape) == (1, 3, h * scale, w * scale)
    psnr, maxabs = gate(got, want)
    print(f"RRDBNet-{blocks} x{scale} {h}x{w}: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0


def test_rrdbnet_tiled_matches_oracle_tiles(engine):
    """tile / tile_pad option (RealESRGANer.tile_process semantics), applied inside the engine plan (crop gather, one
    batch per crop-shape class, paste: csrc/engine.cu create_tiled_plan): compare with the oracle run through the SAME
    tiling (tiled vs untiled cannot meet 50 dB)."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 3, 32).eval()
    x = torch.rand(1, 3, 80, 112)
    with torch.no_grad():
        want = rrdbnet.tile_process(net, x, 2, 48, 10)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=3, device=0, tile=48, tile_pad=10)
    got = model(x.cuda())
    psnr, maxabs = gate(got, want)
    assert psnr >= 50 and maxabs <= 2.0


def test_rrdbnet_tiled_batched_groups(engine):
    """Many tiles per shape class and a batch of two frames: same-shape crops run as one engine batch."""
    torch.manual_seed(1)
    net = rrdbnet.RRDBNet(3, 3, 4, 64, 2, 32).eval()
    x = torch.rand(2, 3, 100, 150)
    with torch.no_grad():
        want = rrdbnet.tile_process(net, x, 4, 32, 6)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=4, num_block=2, device=0, tile=32, tile_pad=6)
    got = model(x.cuda())
    assert tuple(got.shape) == (2, 3, 400, 600)
    psnr, maxabs = gate(got, want)
    assert psnr >= 50 and maxabs <= 2.0


def test_error_behaviour(engine):
    """Errors come back as Ss4kError (a RuntimeError, so BaseService.proc_main's except path fires,
    base_service.py:64-70), never as a crash: CPU tensors, BSVD sizes that are not multiples of 4, missing weights."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 1, 32).eval()
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=1, device=0)
    with pytest.raises(L.Ss4kError):
        model(torch.rand(1, 3, 16, 16))                       # CPU tensor: no CPU path
    sd = {k: v for k, v in net.state_dict().items() if k != "conv_last.weight"}
    with pytest.raises(RuntimeError):
        realesrgan.NativeRRDBNet(sd, scale=2, num_block=1, device=0)(torch.rand(1, 3, 16, 16).cuda())
    from ss4k_b200 import bsvd as nb
    from oracle import bsvd as ob
    den = nb.NativeBSVD(ob.build_bsvd32(0, weight_scale=0.5), device=0)
    with pytest.raises(RuntimeError):
        den(torch.rand(1, 1, 4, 18, 32).cuda())               # H % 4 != 0 (two stride-2 stages)
    with pytest.raises(ValueError):
        den(torch.rand(1, 1, 3, 16, 32).cuda())               # needs RGB + noise map
    # the engine is still usable afterwards
    assert tuple(model(torch.rand(1, 3, 16, 32).cuda()).shape) == (1, 3, 32, 64)


def test_u8_boundary_and_host_path(engine):
    """upscale(frames) boundary: uint8 NHWC in -> uint8 NHWC out, and the host-buffer entry."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    frames = torch.randint(0, 256, (2, 40, 64, 3), dtype=torch.uint8)
    with torch.no_grad():
        want = (net(frames.permute(0, 3, 1, 2) / 255.0).clamp(0, 1) * 255).permute(0, 2, 3, 1)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=2, device=0)
    plan = model._plan(2, 40, 64, L.FMT_U8_NHWC, L.FMT_U8_NHWC)
    got = plan.run(frames.cuda()).cpu()
    assert got.dtype == torch.uint8 and tuple(got.shape) == (2, 80, 128, 3)
    # truncation like the reference (fsrcnn_upscaler.py:233): within 1 LSB + fp16 error of floor(want)
    assert (got.float() - want.floor()).abs().max().item() <= 2
    out_host = torch.empty(2, 80, 128, 3, dtype=torch.uint8).pin_memory()
    plan.run_host(frames.pin_memory(), out_host)
    assert torch.equal(out_host, got)


def test_rrdbnet_uint8_frames_ragged_width(engine):
    """uint8 NHWC frames in / out through the engine plan (the service and bench boundary) at a width that is not a
    multiple of the 32-pixel store rows: packed word stores where a warp's row is complete, byte stores at the ragged
    edge (csrc/conv_stream.cu, last conv)."""
    torch.manual_seed(3)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    frames = torch.randint(0, 256, (2, 44, 150, 3), dtype=torch.uint8)
    with torch.no_grad():
        want = net(frames.permute(0, 3, 1, 2).float() / 255.0).clamp(0, 1)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=2, device=0)
    plan = model._plan(2, 44, 150, L.FMT_U8_NHWC, L.FMT_U8_NHWC)
    got = plan.run(frames.cuda()).cpu()
    assert got.shape == (2, 88, 300, 3) and got.dtype == torch.uint8
    d = (got.permute(0, 3, 1, 2).float() - want * 255.0)
    # truncating quantisation (fsrcnn_upscaler.py:233): got == floor(255 * y) up to the fp16 error of y
    assert d.max().item() <= 0.6 and d.min().item() >= -1.6, (d.min().item(), d.max().item())


def test_rrdbnet_1080p_tile512_every_shape_class(engine):
    """BASELINE.json configs[3] geometry: 1920x1080 frame, tile 512, tile_pad 10 (the reference defaults,
    realesrgan/factory.py:93-94) -> 4 x 3 tiles in 9 padded-crop shape classes (522/532/394 x 522/532/66), run as two
    crop atlases (eight crops side by side in one 3974 x 532 image, the four short ones in a 2000 x 66 image) inside ONE engine run.  RRDBNet x2 with 2 blocks (the tiling, not the depth, is under
    test) against the oracle's tile_process on the whole frame; uint8 frames in and out as the service passes them."""
    torch.manual_seed(2)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    g = torch.Generator().manual_seed(4)
    base = torch.rand(1, 3, 135, 240, generator=g)
    x = (torch.nn.functional.interpolate(base, size=(1080, 1920), mode="bilinear") + 0.03 * torch.randn(1, 3, 1080, 1920, generator=g)).clamp(0, 1)
    with torch.no_grad():
        want = rrdbnet.tile_process(net, x, 2, 512, 10)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=2, device=0, tile=512, tile_pad=10)
    got = model(x.cuda())
    assert tuple(got.shape) == (1, 3, 2160, 3840)
    psnr, maxabs = gate(got, want)
    print(f"RRDBNet-2 x2 1920x1080 tile 512 / pad 10: PSNR {psnr:.1f} dB, max|err| {maxabs:.3f}/255")
    assert psnr >= 50 and maxabs <= 2.0
    plan = model.plan_for(x.cuda())
    assert plan.launches > 2 * 2          # two canvas groups: gather + paste each, plus the nets
    # the same through the uint8 frame boundary
    u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    pu = model._plan(1, 1080, 1920, L.FMT_U8_NHWC, L.FMT_U8_NHWC)
    out = pu.run(u8.cuda()).cpu()
    with torch.no_grad():
        want_u8 = rrdbnet.tile_process(net, u8.permute(0, 3, 1, 2).float() / 255.0, 2, 512, 10).clamp(0, 1) * 255
    d = out.permute(0, 3, 1, 2).float() - want_u8
    assert d.max().item() <= 0.6 and d.min().item() >= -1.6


@pytest.mark.parametrize("h,w,tile,pre_pad", [(45, 71, 0, 0), (44, 70, 0, 6), (45, 71, 32, 5), (60, 90, 40, 4), (45, 70, 32, 0)])
def test_rrdbnet_x2_pre_pad_and_mod_pad(engine, h, w, tile, pre_pad):
    """RealESRGANer.pre_process / post_process: reflect pre_pad on the right / bottom, reflect
    mod-2 pad of the x2 net for odd sizes (upstream pads, it does not refuse), both cropped off the output.  (Tile sizes
    whose padded crops are even: an odd crop fails in upstream's pixel_unshuffle as it does here.)"""
    torch.manual_seed(3)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    x = torch.rand(2, 3, h, w)
    with torch.no_grad():
        want = rrdbnet.enhance_tensor(net, x, 2, tile=tile, tile_pad=6, pre_pad=pre_pad)
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=2, device=0, tile=tile, tile_pad=6, pre_pad=pre_pad)
    got = model(x.cuda())
    assert tuple(got.shape) == (2, 3, 2 * h, 2 * w) == tuple(want.shape)
    psnr, maxabs = gate(got, want)
    assert psnr >= 50 and maxabs <= 2.0


def test_srvgg_x4_tiled_with_pre_pad(engine):
    torch.manual_seed(4)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval()
    x = torch.rand(1, 3, 50, 77)
    with torch.no_grad():
        want = rrdbnet.enhance_tensor(net, x, 4, tile=24, tile_pad=4, pre_pad=3)
    model = realesrgan.NativeSRVGG(net.state_dict(), num_conv=16, upscale=4, device=0, tile=24, tile_pad=4, pre_pad=3)
    got = model(x.cuda())
    assert tuple(got.shape) == (1, 3, 200, 308)
    psnr, maxabs = gate(got, want)
    assert psnr >= 50 and maxabs <= 2.0


@pytest.mark.parametrize("arch", ["rrdb_x2", "rrdb_x4", "srvgg_x4"])
def test_tiled_masked_canvases_equal_exact_classes(engine, arch, monkeypatch):
    """Crops of different shapes share a batch as masked canvases (every conv forces its output back to zero outside the
    image's own crop, so each crop sees the zero padding at its own border): must equal the one-batch-per-shape-class run
    (SS4K_TILE_EXACT_CLASSES=1) bit for bit and launch fewer batches.  Ragged frame, small tiles: nine shape classes."""
    torch.manual_seed(3)
    if arch == "srvgg_x4":
        net = srvgg.SRVGGNetCompact(3, 3, 64, 4, 4).eval()
        mk = lambda: realesrgan.NativeSRVGG(net.state_dict(), num_conv=4, upscale=4, device=0, tile=40, tile_pad=6)  # noqa: E731
    else:
        sc = 2 if arch == "rrdb_x2" else 4
        net = rrdbnet.RRDBNet(3, 3, sc, 64, 1, 32).eval()
        mk = lambda: realesrgan.NativeRRDBNet(net.state_dict(), scale=sc, num_block=1, device=0, tile=40, tile_pad=6)  # noqa: E731
    x = torch.rand(2, 3, 98, 150, generator=torch.Generator().manual_seed(8)).cuda()
    merged = mk()
    a = merged(x).clone()
    n_merged = merged.plan_for(x).launches
    monkeypatch.setenv("SS4K_TILE_EXACT_CLASSES", "1")
    exact = mk()
    b = exact(x)
    assert torch.equal(a, b)
    assert n_merged < exact.plan_for(x).launches
