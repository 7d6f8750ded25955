"""Import the reference's OWN modules by file path (oracle validation; test infrastructure only).

Works only where /root/reference exists (this container, never the GPU box).  Used by
tests/golden/make_golden.py to mint fixtures and by the CPU tests to pin the restatements.
  * SRVGGNetCompact from src/upscale/model/realesrgan/factory.py:18-82 (needs stub ``basicsr`` /
    ``realesrgan`` modules because the file imports them at the top, factory.py:6-9)
  * BSVD from src/upscale/model/bsvd/model.py:467-588 (hard-codes device='cuda' at :87,91,108,123 and
    ``.cuda()`` at :545, so two CPU shims are installed around calls)
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("SS4K_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "src/upscale/model/bsvd/model.py"))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_realesrgan_factory():
    stubs = {
        "basicsr": {}, "basicsr.archs": {}, "basicsr.archs.rrdbnet_arch": {"RRDBNet": object},
        "basicsr.utils": {}, "basicsr.utils.download_util": {"load_file_from_url": lambda **k: None},
        "realesrgan": {"RealESRGANer": object},
    }
    saved = {}
    for k, attrs in stubs.items():
        saved[k] = sys.modules.get(k)
        m = types.ModuleType(k)
        for a, v in attrs.items():
            setattr(m, a, v)
        sys.modules[k] = m
    try:
        return _load("_ref_realesrgan_factory", "src/upscale/model/realesrgan/factory.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_bsvd_model():
    return _load("_ref_bsvd_model", "src/upscale/model/bsvd/model.py")


@contextlib.contextmanager
def cpu_shims():
    """torch.zeros(device='cuda') -> CPU, Tensor.cuda() -> no-op, for the reference BSVD on CPU."""
    real_zeros, real_cuda = torch.zeros, torch.Tensor.cuda

    def zeros(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    torch.zeros = zeros
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.zeros = real_zeros
        torch.Tensor.cuda = real_cuda
