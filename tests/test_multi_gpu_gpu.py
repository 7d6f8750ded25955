"""The multi-GPU service front with REAL per-GPU services (two worker processes; both on cuda:0 when the box has one
GPU): uint8 frames in through the reference's queue interface, upscaled frames back in arrival order, equal to what a
single in-process service returns."""
import pytest
import torch

from ss4k_b200 import multi_gpu, service
from oracle import srvgg

pytestmark = pytest.mark.gpu


def test_two_workers_match_the_single_service(engine):
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval().state_dict().items()}
    kw = dict(lr_level=0, denoising=False, model_name='realesr-animevideov3', state_dict=sd)
    one = service.FsrcnnUpscalerService(device=0, **kw)
    one.proc_init()
    g = torch.Generator().manual_seed(3)
    jobs = [torch.randint(0, 256, (2, 36, 64, 3), dtype=torch.uint8, generator=g) for _ in range(6)]
    want = [one.upscale(f.cuda()).cpu() for f in jobs]
    ndev = torch.cuda.device_count()
    front = multi_gpu.MultiGpuUpscalerService(devices=[0, 1 % ndev], **kw)
    front.start(ready_timeout=600)
    try:
        for i, f in enumerate(jobs):
            front.push_job(service.UpscalerQueueEntry(frames=f.cuda(), step=i, audio_segment=torch.zeros(1)))
        got = [front.get_result(timeout=120) for _ in jobs]
    finally:
        front.stop()
    assert [e.step for e in got] == list(range(len(jobs)))
    for e, w in zip(got, want):
        assert torch.equal(e.frames.cpu(), w)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engines_on_two_gpus_in_one_process(engine):
    """One process, two engines: the second GPU's service, pipeline and BSVD stream run while the calling thread's current
    device is still 0 (the C entry points bind the engine's device for the call and restore the previous one; the Python
    layer does the same around its glue kernels).  Results equal device 0's bit for bit."""
    from ss4k_b200 import _lib as L
    from ss4k_b200 import bsvd as nb, realesrgan
    from ss4k_b200.pipeline import DenoiseUpscalePipeline
    from oracle import bsvd as ob, rrdbnet
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in srvgg.SRVGGNetCompact(3, 3, 64, 16, 4).eval().state_dict().items()}
    kw = dict(lr_level=0, denoising=False, model_name='realesr-animevideov3', state_dict=sd)
    frames = torch.randint(0, 256, (2, 36, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    outs = []
    for dev in (0, 1):
        svc = service.FsrcnnUpscalerService(device=dev, **kw)
        svc.proc_init()
        outs.append(svc.upscale(frames.to(f"cuda:{dev}")).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])
    rr = rrdbnet.RRDBNet(3, 3, 2, 64, 1, 32).eval()
    bsd = ob.build_bsvd32(0)
    nv = torch.randint(16, 236, (20, 48 * 128 * 3 // 2), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    res, streams = [], []
    for dev in (0, 1):
        den = nb.NativeBSVD(bsd, device=dev, act_mode=L.ACT_F16_SPLIT, out_dtype=torch.float16)
        sr = realesrgan.NativeRRDBNet(rr.state_dict(), scale=2, num_block=1, device=dev)
        pipe = DenoiseUpscalePipeline(den, sr, 48, 128, 0.075, nv12=True)
        res.append(pipe.run(nv.to(f"cuda:{dev}"), slice(2, 6)).cpu())
        s = den.stream(48, 128, in_fmt=L.FMT_NV12, noise=0.075)
        x = nv.to(f"cuda:{dev}")
        got = [o for o in (s.push(x[i].reshape(72, 128)) for i in range(20)) if o is not None] + list(s.flush())
        s.close()
        streams.append(torch.cat(got).cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(res[0], res[1]) and torch.equal(streams[0], streams[1])
