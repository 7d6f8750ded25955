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


def load_fsrcnn_service_code():
    """The reference's OWN service glue, compiled from its source text: ``blur_ker``, ``sharpen_ker`` and the methods
    ``upscale`` / ``upscale_multi`` / ``upscale_single`` of ``FsrcnnUpscalerService``
    (src/upscale/fsrcnn_upscaler.py:20-84,144-326).

    The module itself cannot be imported (top-level ``matplotlib`` / ``tqdm`` and relative imports of all three model
    factories), so the function and method definitions are cut out of its AST unchanged and
    executed in a namespace that holds what they use (torch, F, math, time).  Returns that namespace; the class has
    no base and no ``__init__`` -- build instances with ``object.__new__`` and set the attributes the methods read."""
    import ast
    import math
    import time
    with open(os.path.join(REF_ROOT, "src/upscale/fsrcnn_upscaler.py")) as f:
        tree = ast.parse(f.read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("blur_ker", "sharpen_ker"):
            keep.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "FsrcnnUpscalerService":
            node.bases, node.keywords, node.decorator_list = [], [], []
            node.body = [n for n in node.body if isinstance(n, ast.FunctionDef) and
                         n.name in ("upscale", "upscale_multi", "upscale_single")]
            keep.append(node)
    assert len(keep) == 3, [getattr(k, "name", None) for k in keep]
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"torch": torch, "F": torch.nn.functional, "math": math, "time": time}
    exec(compile(mod, os.path.join(REF_ROOT, "src/upscale/fsrcnn_upscaler.py"), "exec"), ns)
    return ns


class _NullProfiler:
    def start(self, name):
        pass

    def end(self, name):
        pass


def make_reference_service(ns, model, lr_shape, output_shape=None, lr_hr_resize=True, denoise_model=None,
                           denoise_rate=1.0):
    """An instance of the reference's FsrcnnUpscalerService (see load_fsrcnn_service_code) wired like proc_init
    (fsrcnn_upscaler.py:118-139) but on the CPU in fp32: the stencil modules are the reference's blur_ker / sharpen_ker
    without the ``.half()`` (torch.cuda.amp.autocast is a no-op without a GPU), the nets are the callables given."""
    svc = object.__new__(ns["FsrcnnUpscalerService"])
    svc.lr_shape, svc.output_shape, svc.lr_hr_resize = tuple(lr_shape), output_shape, lr_hr_resize
    svc.device = torch.device("cpu")
    svc.upscaler_model = "realesrgan"
    svc.single_mode = False
    svc.denoising = denoise_model is not None
    svc.denoise_rate = denoise_rate
    svc.profiler = _NullProfiler()
    svc.model = model
    svc.lr_prev = None
    svc.match_blur = ns["blur_ker"](kernel_size=8 * 2 + 1, sigma=8.0)
    if denoise_model is not None:
        svc.denoise_model = denoise_model
        svc.denoise_blur = ns["blur_ker"]()
        svc.denoise_sharpen = ns["sharpen_ker"](strength=0.00002)
        svc.denoise_sharpen_hr = ns["sharpen_ker"](strength=0.00007)
    return svc
