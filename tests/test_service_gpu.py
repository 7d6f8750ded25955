"""GPU parity of the drop-in upscaler service (uint8 frames in -> uint8 frames out) against the oracle glue
(oracle/glue.py, a line-by-line fp32 restatement of fsrcnn_upscaler.py:168-326) on the same seeded weights."""
import pytest
import torch

from ss4k_b200 import service
from oracle import bsvd, glue, rrdbnet, srvgg

pytestmark = pytest.mark.gpu


def _cmp(got, want, max_lsb=3, mean_lsb=0.6):
    """``want`` = (uint8 frame, float frame x 255 in front of the truncating cast) from the oracle glue.
    The north-star gate is on the FLOAT frame: |engine - oracle| <= 2/255.  The engine truncates like the reference
    (fsrcnn_upscaler.py:233), so its uint8 value g = floor(q_engine) lies in (q_oracle - 3, q_oracle + 2]: that interval is
    what is asserted.  On the uint8 frames it allows 3 LSB (|floor(a) - floor(b)| <= floor(|a - b|) + 1), the third LSB
    being the truncation flip of a value next to an integer boundary, never a larger float error."""
    want_u8, want_q = want
    assert got.dtype == torch.uint8 and got.shape == want_u8.shape, (got.shape, want_u8.shape)
    g = got.cpu()
    d = (g.int() - want_u8.int()).abs()
    e = g.float() - want_q
    print(f"uint8 diff: mean {d.float().mean().item():.3f}, max {d.max().item()}, >1 LSB {100.0 * (d > 1).float().mean().item():.3f}%; "
          f"uint8 - oracle float: [{e.min().item():.3f}, {e.max().item():.3f}]")
    assert e.max().item() <= 2.0 and e.min().item() > -3.0
    assert d.max().item() <= max_lsb and d.float().mean().item() <= mean_lsb


def _frames(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n, h // 8 + 1, w // 8 + 1, 3, generator=g)
    up = torch.nn.functional.interpolate(base.permute(0, 3, 1, 2), size=(h, w), mode="bilinear").permute(0, 2, 3, 1)
    return (up * 255 + torch.randn(n, h, w, 3, generator=g) * 6).clamp(0, 255).to(torch.uint8)


def test_upscale_multi_rrdb_x2(engine):
    """Default wiring (upscale_multi): RRDBNet x2, colour-distribution + local colour match, uint8 out."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    frames = _frames(2, 96, 160, 1)
    want = glue.upscale_multi(frames, net, lr_shape=(720, 1280), output_shape=(192, 320), return_float=True)
    svc = service.FsrcnnUpscalerService(lr_level=3, device=0, denoising=False, model_name='RealESRGAN_x2plus',
                                        state_dict=net.state_dict(), batch_size=2)
    svc.proc_init()
    svc.output_shape = (192, 320)
    got = svc.upscale(frames.cuda())
    torch.cuda.synchronize()
    _cmp(got, want)


def test_upscale_multi_srvgg_x4_bicubic_down(engine):
    """The reference's live default: SRVGG x4 then bicubic resize to output_shape (fsrcnn_upscaler.py:222-231),
    with the LR area-downscale branch (:173-176) because the input is larger than lr_shape."""
    torch.manual_seed(0)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval()
    frames = _frames(1, 2 * 360, 2 * 640, 2)
    want = glue.upscale_multi(frames, net, lr_shape=(360, 640), output_shape=(720, 1280), return_float=True)
    svc = service.FsrcnnUpscalerService(lr_level=0, device=0, denoising=False, model_name='realesr-animevideov3',
                                        state_dict=net.state_dict())
    svc.proc_init()
    svc.output_shape = (720, 1280)
    got = svc.upscale(frames.cuda())
    _cmp(got, want)


@pytest.mark.parametrize("out_shape", [(225, 400), (300, 530), (400, 700), (180, 320), (100, 200)])
def test_upscale_multi_resize_factors(engine, out_shape):
    """Bicubic resize to output_shape (fsrcnn_upscaler.py:222-231) at non-integer factors, up- and down-scaling: the
    fused finalise + bicubic tile kernel (factors <= 2) and the two-pass path (factor > 2) against the oracle."""
    torch.manual_seed(0)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval()
    frames = _frames(2, 90, 160, 5)
    want = glue.upscale_multi(frames, net, lr_shape=(720, 1280), output_shape=out_shape, return_float=True)
    svc = service.FsrcnnUpscalerService(lr_level=3, device=0, denoising=False, model_name='realesr-animevideov3',
                                        state_dict=net.state_dict(), batch_size=2)
    svc.proc_init()
    svc.output_shape = out_shape
    got = svc.upscale(frames.cuda())
    assert tuple(got.shape) == (2,) + tuple(out_shape) + (3,)
    _cmp(got, want)


def test_upscale_single_denoise_then_rrdb(engine):
    """The composition the north star names (dead by default in the reference): BSVD denoise
    (F = 1 clip, noise map 0.05 on the first frame) -> sharpen/blend -> RRDBNet x2 -> HR sharpen -> match."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 2, 32).eval()
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    frames = _frames(2, 360, 640, 3)
    den = lambda x: bsvd.bsvd_forward(sd, x)  # noqa: E731
    want0 = glue.upscale_single(frames[0], net, (360, 640), (720, 1280), den, 0.75, True, return_float=True)
    want1 = glue.upscale_single(frames[1], net, (360, 640), (720, 1280), den, 0.75, False, return_float=True)
    svc = service.FsrcnnUpscalerService(lr_level=0, device=0, denoising=True, denoise_rate=0.75,
                                        model_name='RealESRGAN_x2plus', state_dict=net.state_dict(),
                                        denoise_state_dict=sd, single_mode=True)
    svc.proc_init()
    svc.output_shape = (720, 1280)
    got = svc.upscale(frames.cuda())
    _cmp(got[0], want0)
    _cmp(got[1], want1)


def test_upscale_temporal_ring_buffers(engine):
    """temporal_denoise=True: upscale_single fed through BSVD's persistent ring buffers (bsvd/model.py:510-513) -- every
    frame denoised with its temporal neighbours, jobs delayed as a whole by the 16-frame pipeline, flush() at the end.
    Oracle: BSVD.forward over the WHOLE clip (== streaming_forward, model.py:515-580) on the per-frame inputs the service
    builds (noise map 0.05 on the first frame), then the reference's per-frame glue on each denoised frame."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 1, 32).eval()
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    t, lh, lw = 18, 360, 640
    frames = _frames(t, lh, lw, 9)
    x = torch.empty(1, t, 4, lh, lw)
    x[0, :, :3] = frames.permute(0, 3, 1, 2) / 255.0        # lr_shape == frame size: the area resize is the identity
    x[0, :, 3] = 0.075
    x[0, 0, 3] = 0.05
    clip = bsvd.bsvd_forward(sd, x)
    svc = service.FsrcnnUpscalerService(lr_level=0, device=0, denoising=True, denoise_rate=0.75, batch_size=4,
                                        model_name='RealESRGAN_x2plus', state_dict=net.state_dict(),
                                        denoise_state_dict=sd, single_mode=True, temporal_denoise=True)
    svc.proc_init()
    svc.output_shape = (720, 1280)
    entries = []
    jobs = [(k, frames[4 * k: 4 * k + 4]) for k in range((t + 3) // 4)]
    for k, fr in jobs:
        e = svc.proc_job_recieved(service.UpscalerQueueEntry(frames=fr.cuda(), step=100 + k, audio_segment=("audio", k)))
        if k < 4:   # 16 frames of latency: nothing can leave before the 17th push (job 4)
            assert e.frames.shape[0] == 0 and e.step == -1 and e.audio_segment is None
        entries.append(e)
    entries += svc.flush()
    done = [e for e in entries if e.frames.shape[0]]
    assert [e.step for e in done] == [100 + k for k, _ in jobs]                 # in order, step / audio of the delayed job
    assert [e.audio_segment for e in done] == [("audio", k) for k, _ in jobs]
    assert [e.frames.shape[0] for e in done] == [fr.shape[0] for _, fr in jobs]
    got = torch.cat([e.frames for e in done], dim=0)
    assert tuple(got.shape) == (t, 720, 1280, 3)
    for i in (0, 7, t - 1):
        want = glue.upscale_single(frames[i], net, (lh, lw), (720, 1280), lambda _x, i=i: clip[:, i:i + 1], 0.75, i == 0,
                                   return_float=True)
        _cmp(got[i], want)
    # a frame inside the stream differs from the same frame denoised alone (the F = 1 path of the reference)
    alone = service.FsrcnnUpscalerService(lr_level=0, device=0, denoising=True, denoise_rate=0.75,
                                          model_name='RealESRGAN_x2plus', state_dict=net.state_dict(),
                                          denoise_state_dict=sd, single_mode=True)
    alone.proc_init()
    alone.output_shape = (720, 1280)
    alone.lr_prev = 1     # not the first frame: noise map 0.075
    assert (alone.upscale(frames[7:8].cuda())[0].int() - got[7].int()).abs().max().item() > 0
    # a new clip after flush(): same result as the first time
    e = [svc.proc_job_recieved(service.UpscalerQueueEntry(frames=fr.cuda(), step=k)) for k, fr in jobs]
    got2 = torch.cat([q.frames for q in e + svc.flush() if q.frames.shape[0]], dim=0)
    assert torch.equal(got2, got)
    svc.proc_cleanup()


def test_proc_job_recieved_roundtrip(engine):
    torch.manual_seed(0)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval()
    svc = service.FsrcnnUpscalerService(lr_level=0, device=0, denoising=False, model_name='realesr-animevideov3',
                                        state_dict=net.state_dict())
    svc.proc_init()
    job = service.UpscalerQueueEntry(frames=_frames(1, 40, 72, 4).cuda(), step=7)
    out = svc.proc_job_recieved(job)
    assert out.step == 7 and tuple(out.frames.shape) == (1, 160, 288, 3) and out.frames.dtype == torch.uint8


def test_image_server_shapes_lru(engine):
    """The still-image server drives the same service with arbitrary image sizes (image_pipeline.py:54-64:
    lr_level=3, batch_size=1, lr_hr_resize=False; one job per image): one engine plan per shape, the least recently
    used plan is destroyed once max_plans are cached, and a re-planned shape reproduces its first result bit for bit."""
    torch.manual_seed(0)
    net = srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval()
    svc = service.FsrcnnUpscalerService(lr_level=3, device=0, denoising=False, model_name='realesr-animevideov3',
                                        state_dict=net.state_dict(), batch_size=1, lr_hr_resize=False)
    svc.proc_init()
    svc.model._plans.max_plans = 2
    shapes = [(40, 56), (72, 96), (33, 130), (40, 56)]
    outs = []
    for i, (h, w) in enumerate(shapes):
        frames = _frames(1, h, w, 10 + (i % 3))
        got = svc.upscale(frames.cuda())
        torch.cuda.synchronize()
        assert got.shape == (1, 4 * h, 4 * w, 3)
        want = glue.upscale_multi(frames, net, lr_shape=(720, 1280), output_shape=None, lr_hr_resize=False, return_float=True)
        _cmp(got, want)
        outs.append(got.cpu())
    assert svc.model._plans.evictions >= 2 and len(svc.model._plans) == 2
    assert torch.equal(outs[0], outs[3])
