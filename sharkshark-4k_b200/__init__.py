"""sharkshark-4k hot path, B200-native: BSVD denoiser -> RealESRGAN upscaler behind the reference's
``src/upscale`` model / service interface.  All arithmetic runs in ``csrc/libss4k.so`` (hand-written
sm_100a CUDA: tcgen05 / TMEM / TMA); importing this package never imports ``oracle``.

The directory name carries a hyphen (it is the name the task prescribes), so import it with
``importlib.import_module("sharkshark-4k_b200")`` or through the root-level alias ``ss4k_b200``.
"""
from . import _lib
from ._lib import Ss4kError  # noqa: F401
from .engine import Engine, Plan, make_cfg, plan_dry  # noqa: F401

__all__ = ["Engine", "Plan", "Ss4kError", "make_cfg", "plan_dry"]
