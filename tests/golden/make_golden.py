"""Mint the golden fixtures from the REFERENCE's own modules (run in the build container only:
needs /root/reference).  Usage:  python tests/golden/make_golden.py

The reference pins no numeric result of the hot path, so the fixtures are
outputs of the reference nn.Modules themselves, imported by file path:
  srvgg16_x4.npz : SRVGGNetCompact(3,3,64,16,4,'prelu')  src/upscale/model/realesrgan/factory.py:18-82
  srvgg32_x4.npz : SRVGGNetCompact(3,3,64,32,4,'prelu')  (the service default, factory.py:132-134)
  bsvd32_f5.npz  : BSVD(chns=[32,64,128], mid_ch=32, interm_ch=30, relu6, norm none).forward on a
                   5-frame clip   src/upscale/model/bsvd/model.py:467-588 (config bsvd/factory.py:31-35)
  bsvd32_f1.npz  : the same net on a 1-frame clip (what the service feeds, fsrcnn_upscaler.py:277)
  glue.npz       : uint8 outputs of the reference's own FsrcnnUpscalerService.upscale_multi / upscale_single code
                   (fsrcnn_upscaler.py:168-326, compiled from its source text) around weight-free stand-in nets
Each file holds the seeded input, the reference fp32 CPU output and per-tensor weight checksums
(sum, sum of squares) so that tests can rebuild the weights from the seed through oracle/ and prove
they are the reference's weights without shipping megabytes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_import as ri  # noqa: E402


GLUE_MULTI_CASES = (((720, 1280), None), ((48, 64), (100, 140)), ((48, 64), None))
GLUE_SINGLE_LR, GLUE_SINGLE_OUT = (48, 64), (100, 140)


def glue_models():
    """Stand-in nets for the glue fixtures: an x2 'upscaler' and a 'denoiser' that need no weights."""
    import torch.nn.functional as F
    model = lambda t: F.interpolate(t.float(), scale_factor=2, mode="bicubic", align_corners=False) * 0.9 + 0.04  # noqa: E731
    den = lambda x: x[:, :, :3] * 0.97 + 0.01  # noqa: E731
    return model, den


def glue_frames():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(7)
    base = torch.rand(2, 13, 17, 3, generator=g)
    smooth = F.interpolate(base.permute(0, 3, 1, 2), size=(96, 128), mode="bilinear").permute(0, 2, 3, 1)
    return (smooth * 255 + torch.randn(2, 96, 128, 3, generator=g) * 6).clamp(0, 255).to(torch.uint8)


def checksums(sd):
    keys = sorted(sd.keys())
    arr = np.array([[sd[k].double().sum().item(), (sd[k].double() ** 2).sum().item()] for k in keys])
    return np.array(keys), arr


def main():
    assert ri.available(), "needs /root/reference"
    torch.set_num_threads(1)  # one accumulation order
    fac = ri.load_realesrgan_factory()
    for nconv in (16, 32):
        torch.manual_seed(0)
        net = fac.SRVGGNetCompact(3, 3, 64, nconv, 4, "prelu").eval()
        x = torch.rand(1, 3, 12, 20, generator=torch.Generator().manual_seed(1234))
        with torch.no_grad():
            y = net(x)
        k, c = checksums(net.state_dict())
        np.savez_compressed(os.path.join(HERE, f"srvgg{nconv}_x4.npz"), x=x.numpy(), y=y.numpy(), keys=k, sums=c,
                            seed=0)
    bm = ri.load_bsvd_model()
    for frames in (5, 1):
        with ri.cpu_shims():
            torch.manual_seed(0)
            net = bm.BSVD(chns=[32, 64, 128], mid_ch=32, shift_input=False, norm="none", interm_ch=30, act="relu6",
                          pretrain_ckpt=None).eval()
            x = torch.rand(1, frames, 4, 16, 24, generator=torch.Generator().manual_seed(1234))
            x[:, :, 3] = 0.075
            with torch.no_grad():
                y = net(x)
        k, c = checksums(net.state_dict())
        np.savez_compressed(os.path.join(HERE, f"bsvd32_f{frames}.npz"), x=x.numpy(), y=y.numpy(), keys=k, sums=c,
                            seed=0)
    # service glue: the reference's OWN upscale_multi / upscale_single code (fsrcnn_upscaler.py:168-326), cut out of its
    # source (oracle/reference_import.py::load_fsrcnn_service_code), run on the CPU in fp32 around stand-in nets that
    # tests can rebuild without weights (glue_models below)
    import warnings
    warnings.simplefilter("ignore")  # torch.cuda.amp.autocast without a GPU
    ns = ri.load_fsrcnn_service_code()
    model, den = glue_models()
    frames = glue_frames()
    out = {"frames": frames.numpy()}
    for i, (lr_shape, out_shape) in enumerate(GLUE_MULTI_CASES):
        svc = ri.make_reference_service(ns, model, lr_shape, out_shape)
        out[f"multi{i}"] = svc.upscale_multi(frames.clone()).numpy()
    for i, use_den in enumerate((False, True)):
        svc = ri.make_reference_service(ns, model, GLUE_SINGLE_LR, GLUE_SINGLE_OUT, denoise_model=den if use_den else None,
                                        denoise_rate=0.75)
        for fi in range(2):   # two consecutive frames: first-frame and steady-state noise maps (:262,269)
            out[f"single{i}_{fi}"] = svc.upscale_single(frames[fi].clone()).numpy()
    np.savez_compressed(os.path.join(HERE, "glue.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
