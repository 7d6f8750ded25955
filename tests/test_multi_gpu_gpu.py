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
