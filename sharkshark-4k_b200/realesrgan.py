"""Drop-in for the reference's RealESRGAN model factory
(reference: src/upscale/model/realesrgan/factory.py:85-98 ArgsData, :108-234 build_model, :236-245 JitWrapper).

``build_model(factor, device, input_shape, batch_size, denoise_rate, jit_mode)`` keeps the reference
signature and returns an ``nn.Module`` whose ``forward(x)`` takes ``[N,3,H,W]`` in [0,1] on ``device`` and
returns ``[N,3,s*H,s*W]`` (un-clamped), exactly what ``FsrcnnUpscalerService`` calls
(src/upscale/fsrcnn_upscaler.py:181,294).  The arithmetic runs in libss4k.so (tcgen05 convs with fused
epilogues); plans are built lazily per input shape, so arbitrary shapes work (image server path,
src/sharkshark/image_server/image_pipeline.py:54-64).
"""
import os
from dataclasses import dataclass

import torch
from torch import nn

from . import _lib as L
from .engine import Engine, PlanCache


@dataclass
class ArgsData:
    # same defaults as the reference (factory.py:85-98)
    model_name = 'realesr-general-x4v3'
    denoise_strength = 0.5
    outscale = 4
    model_path = None
    tile = 0
    tile_pad = 10
    pre_pad = 0


# model_name -> (arch, scale, depth)   (factory.py:112-138)
MODEL_ZOO = {
    'RealESRGAN_x4plus': (L.ARCH_RRDB, 4, 23),
    'RealESRNet_x4plus': (L.ARCH_RRDB, 4, 23),
    'RealESRGAN_x4plus_anime_6B': (L.ARCH_RRDB, 4, 6),
    'RealESRGAN_x2plus': (L.ARCH_RRDB, 2, 23),
    'realesr-animevideov3': (L.ARCH_SRVGG, 4, 16),
    'realesr-general-x4v3': (L.ARCH_SRVGG, 4, 32),
}


def dni(net_a, net_b, dni_weight):
    """RealESRGANer.dni (realesrgan/utils.py @5ca1078): w = dni_weight[0]*w_a + dni_weight[1]*w_b for
    every key; the reference passes [denoise_strength, 1 - denoise_strength] (factory.py:152-157)."""
    return {k: dni_weight[0] * v + dni_weight[1] * net_b[k] for k, v in net_a.items()}


def load_checkpoint(path):
    """RealESRGANer weight load: key 'params_ema' if present else 'params'."""
    ck = torch.load(path, map_location='cpu')
    if 'params_ema' in ck:
        return ck['params_ema']
    if 'params' in ck:
        return ck['params']
    return ck


class _NativeNet(nn.Module):
    """nn.Module facade over engine plans (``.eval()``, ``.to()``, ``.half()`` are accepted and ignored:
    weights live inside the engine in its packed 16-bit layout)."""

    arch = None

    def __init__(self, state_dict, scale, depth, device=0, act_mode=L.ACT_F16, tile=0, tile_pad=10, pre_pad=0,
                 out_dtype=torch.float32, use_graph=True, max_plans=12):
        super().__init__()
        self.engine = Engine.get(device)
        self.scale, self.depth, self.act_mode = scale, depth, act_mode
        # RealESRGANer options (factory.py:93-95): applied inside the engine plan (csrc/engine.cu create_tiled_plan)
        self.tile, self.tile_pad, self.pre_pad = tile, tile_pad, pre_pad
        self.out_dtype = out_dtype
        self.use_graph = use_graph
        self.net_id = self.engine.new_net(state_dict)
        self._plans = PlanCache(max_plans)

    def _plan(self, n, h, w, in_fmt, out_fmt):
        key = (n, h, w, in_fmt, out_fmt, self.tile, self.tile_pad, self.pre_pad)
        return self._plans.get(key, lambda: self.engine.plan(
            self.net_id, self.arch, n, h, w, scale=self.scale, depth=self.depth, tile=self.tile or 0, tile_pad=self.tile_pad,
            act_mode=self.act_mode, in_fmt=in_fmt, out_fmt=out_fmt, use_graph=self.use_graph, pre_pad=self.pre_pad or 0))

    def plan_for(self, x):
        n, _, h, w = x.shape
        in_fmt = L.FMT_F16_NCHW if x.dtype == torch.float16 else L.FMT_F32_NCHW
        out_fmt = L.FMT_F16_NCHW if self.out_dtype == torch.float16 else L.FMT_F32_NCHW
        return self._plan(n, h, w, in_fmt, out_fmt)

    def forward(self, x):
        if not x.is_cuda:
            raise L.Ss4kError("native model called with a CPU tensor: there is no CPU path")
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        x = x.contiguous()
        return self.plan_for(x).run(x)

    # weights are not nn.Parameters: these keep callers such as ``model.eval().to(device)`` working
    def half(self):
        return self

    def float(self):
        return self

    def close(self):
        """Destroy the cached plans and release the engine's host copy of the weights (ss4k_clear_weights)."""
        for p in list(self._plans._d.values()):
            p.close()
        self._plans._d.clear()
        if self.net_id is not None:
            self.engine.release_net(self.net_id)
            self.net_id = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeSRVGG(_NativeNet):
    """SRVGGNetCompact (factory.py:18-82) on the native engine."""
    arch = L.ARCH_SRVGG

    def __init__(self, state_dict, num_conv=16, upscale=4, device=0, act_mode=L.ACT_F16, **kw):
        super().__init__(state_dict, upscale, num_conv, device=device, act_mode=act_mode, **kw)


class NativeRRDBNet(_NativeNet):
    """basicsr RRDBNet (constructors at factory.py:113-125) on the native engine."""
    arch = L.ARCH_RRDB

    def __init__(self, state_dict, scale=4, num_block=23, device=0, act_mode=L.ACT_F16, **kw):
        super().__init__(state_dict, scale, num_block, device=device, act_mode=act_mode, **kw)


def build_model(factor=4, device=0, input_shape=(720, 1280), batch_size=8, denoise_rate=0.5, jit_mode=None,
                args=None, state_dict=None, act_mode=L.ACT_F16):
    """Same signature as the reference's build_model (factory.py:108).  ``jit_mode`` is accepted for
    compatibility; every mode maps to the native engine ('b200').  ``state_dict`` (or
    ``args.model_path``) supplies the weights; there is no network here, so nothing is downloaded."""
    args = args or ArgsData()
    args.denoise_strength = denoise_rate
    name = args.model_name.split('.')[0]
    if name not in MODEL_ZOO:
        raise ValueError(f"unknown model_name {name}")
    arch, netscale, depth = MODEL_ZOO[name]
    # DNI (factory.py:152-157): 'realesr-general-x4v3' with denoise_strength != 1 ALWAYS blends the general and the
    # wdn weights in the reference; a missing second weight set is an error here, never a silent un-blended net.
    needs_dni = name == 'realesr-general-x4v3' and args.denoise_strength != 1
    dni_weight = [args.denoise_strength, 1 - args.denoise_strength]
    if state_dict is None:
        path = args.model_path or os.path.join('weights', name + '.pth')
        if isinstance(path, (list, tuple)):
            sds = [load_checkpoint(p) for p in path]
            state_dict = dni(sds[0], sds[1], dni_weight)
        else:
            if not os.path.isfile(path):
                raise FileNotFoundError(f"{path}: weights are downloaded at run time by the reference "
                                        "(factory.py:140-150); pass state_dict= or args.model_path")
            state_dict = load_checkpoint(path)
            if needs_dni:
                wdn = path.replace('realesr-general-x4v3', 'realesr-general-wdn-x4v3')
                if wdn == path or not os.path.isfile(wdn):
                    raise FileNotFoundError(
                        f"{wdn}: model 'realesr-general-x4v3' with denoise_strength {args.denoise_strength} != 1 blends "
                        "the general and the wdn weights (factory.py:152-157); the wdn checkpoint is missing")
                state_dict = dni(state_dict, load_checkpoint(wdn), dni_weight)
    elif isinstance(state_dict, (list, tuple)):
        if len(state_dict) != 2:
            raise ValueError("state_dict=(general, wdn): exactly two weight sets")
        state_dict = dni(state_dict[0], state_dict[1], dni_weight)
    elif needs_dni:
        raise ValueError(
            f"model 'realesr-general-x4v3' with denoise_strength {args.denoise_strength} != 1 needs both weight sets: "
            "pass state_dict=(general_state_dict, wdn_state_dict) (the reference always blends them, factory.py:152-157)")
    # depth follows the weights actually supplied (the zoo entry is the published architecture)
    if arch == L.ARCH_RRDB:
        blocks = {int(k.split('.')[1]) for k in state_dict if k.startswith('body.') and '.rdb1.conv1.weight' in k}
        depth = max(blocks) + 1 if blocks else depth
    else:
        convs = [k for k, v in state_dict.items() if k.startswith('body.') and k.endswith('.weight') and v.ndim == 4]
        depth = len(convs) - 2 if len(convs) >= 2 else depth
    cls = NativeSRVGG if arch == L.ARCH_SRVGG else NativeRRDBNet
    kw = dict(device=device, act_mode=act_mode, tile=args.tile, tile_pad=args.tile_pad, pre_pad=args.pre_pad)
    if arch == L.ARCH_SRVGG:
        model = cls(state_dict, num_conv=depth, upscale=netscale, **kw)
    else:
        model = cls(state_dict, scale=netscale, num_block=depth, **kw)
    model.netscale = netscale
    return model.eval()
