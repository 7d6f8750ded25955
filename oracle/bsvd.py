"""BSVD restated as a whole-clip function (oracle; test infrastructure only).

The reference streams frames through 16 stateful BiBufferConv layers and three FIFO skips
(/root/reference/src/upscale/model/bsvd/model.py:59-138,332-350,402-424,526-580).  For a clip of T
frames that is equivalent to the frame-aligned formulation below:
  shift conv (model.py:43-52,94-138):  out_t = conv(cat[X_{t+1}[:, :C/8], X_{t-1}[:, C/8:C/4], X_t[:, C/4:]])
      with zero features for t+1 >= T and t-1 < 0 (clip start :105-108, end :123);
  DenBlock (model.py:402-424): inc -> downc0 -> downc1 -> upc2 -> (+skip3) upc1 -> (+skip2) outc ->
      out[:, :3] = in[:, :3] - out[:, :3]   (:436-442)
  BSVD.feedin_one_element (model.py:510-513): temp2(temp1(x)).
Equivalence with the streaming reference is checked bit-exactly in tests/test_oracle_cpu.py.
Config used by the service (bsvd/factory.py:31-35): chns=[32,64,128], mid_ch=32, interm_ch=30,
act='relu6', norm='none', in_ch=4, out_ch=3.
The functions take the reference module's own state_dict (same key names).
"""
import torch
from torch.nn import functional as F


def _conv(sd, key, x, stride=1):
    # frame by frame, like the streaming reference: a batched conv2d picks another accumulation
    # order on the CPU (1.4e-5 abs difference), per-frame calls reproduce the reference bit-exactly
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    return torch.cat([F.conv2d(x[i:i + 1], w, b, stride=stride, padding=1) for i in range(x.shape[0])])


def shift_conv(sd, key, x):
    """x: [T, C, H, W] (time on dim 0)."""
    t, c, h, w = x.shape
    fold = c // 8
    nxt = torch.zeros_like(x[:, :fold])
    nxt[:-1] = x[1:, :fold]                     # X_{t+1}[:fold], zero at the clip end
    prv = torch.zeros_like(x[:, fold:2 * fold])
    prv[1:] = x[:-1, fold:2 * fold]             # X_{t-1}[fold:2fold], zero at the clip start
    return _conv(sd, key + ".op.conv", torch.cat([nxt, prv, x[:, 2 * fold:]], dim=1))


def relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def mem_cv_block(sd, key, x):
    x = relu6(shift_conv(sd, key + ".c1", x))
    return relu6(shift_conv(sd, key + ".c2", x))


def den_block(sd, p, x):
    x0 = relu6(_conv(sd, p + "inc.convblock.0", x))
    x0 = relu6(_conv(sd, p + "inc.convblock.3", x0))
    x1 = relu6(_conv(sd, p + "downc0.convblock.0", x0, stride=2))
    x1 = mem_cv_block(sd, p + "downc0.memconv", x1)
    x2 = relu6(_conv(sd, p + "downc1.convblock.0", x1, stride=2))
    x2 = mem_cv_block(sd, p + "downc1.memconv", x2)
    x2 = mem_cv_block(sd, p + "upc2.memconv", x2)
    x2 = F.pixel_shuffle(_conv(sd, p + "upc2.convblock.0", x2), 2)
    x1 = mem_cv_block(sd, p + "upc1.memconv", x2 + x1)
    x1 = F.pixel_shuffle(_conv(sd, p + "upc1.convblock.0", x1), 2)
    y = relu6(_conv(sd, p + "outc.convblock.0", x1 + x0))
    y = _conv(sd, p + "outc.convblock.3", y)
    y = y.clone()
    y[:, :3] = x[:, :3] - y[:, :3]
    return y


def bsvd_clip(sd, x):
    """x: [T, 4, H, W] noisy RGB + noise map  ->  [T, 3, H, W]."""
    with torch.no_grad():
        return den_block(sd, "temp2.", den_block(sd, "temp1.", x))


def bsvd_forward(sd, inp):
    """Mirror of BSVD.forward (model.py:515-524): inp [N, F, 4, H, W] -> [N, F, 3, H, W].
    NOTE the reference reshapes N*F into ONE stream, so clips of a batch are concatenated in time."""
    n, f, c, h, w = inp.shape
    out = bsvd_clip(sd, inp.reshape(n * f, c, h, w))
    return out.reshape(n, f, 3, h, w)


# --------------------------------------------------------------------------------------------------
# Weight construction.  /root/reference does not exist on the GPU box, so the oracle must be able to
# mint the reference's *random-init* BSVD weights by itself.  The classes below restate ONLY the
# constructors of the reference modules (same attribute names -> same state-dict keys, same creation
# order and the same three init passes -> same RNG consumption), so ``build_bsvd32(seed)`` reproduces
# ``torch.manual_seed(seed); BSVD(chns=[32,64,128], mid_ch=32, interm_ch=30, act='relu6',
# norm='none', pretrain_ckpt=None)`` bit for bit (checked against the reference in
# tests/test_oracle_cpu.py and pinned by the checksums in tests/golden/).
#   constructors: bsvd/model.py:22-41 ShiftConv, :63-91 BiBufferConv, :143-154 MemCvBlock,
#   :221-240 InputCvBlock, :256-267 DownBlock, :285-292 UpBlock, :314-323 OutputCvBlock,
#   :362-381 DenBlock (+ reset_params :393-400), :471-481 BSVD (+ reset_params :501-508)
from torch import nn  # noqa: E402


def _conv3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=3, padding=1, stride=stride, bias=True)


class _ShiftConvParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = _conv3(c, c)


class _BiBufferParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.op = _ShiftConvParams(c)


class _MemCvParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.c1 = _BiBufferParams(c)
        self.c2 = _BiBufferParams(c)


class _Seq(nn.Module):
    """Holds convs under the numeric child names nn.Sequential would give them."""

    def __init__(self, convs):
        super().__init__()
        for idx, m in convs:
            self.add_module(str(idx), m)


class _InputParams(nn.Module):
    def __init__(self, in_ch, interm, out_ch):
        super().__init__()
        self.convblock = _Seq([(0, _conv3(in_ch, interm)), (3, _conv3(interm, out_ch))])


class _DownParams(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.convblock = _Seq([(0, _conv3(cin, cout, stride=2))])
        self.memconv = _MemCvParams(cout)


class _UpParams(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.memconv = _MemCvParams(cin)
        self.convblock = _Seq([(0, _conv3(cin, cout * 4))])


class _OutParams(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.convblock = _Seq([(0, _conv3(cin, cin)), (3, _conv3(cin, cout))])


def _kaiming_all(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, nonlinearity="relu")


class _DenBlockParams(nn.Module):
    def __init__(self, chns, out_ch, in_ch, interm_ch):
        super().__init__()
        c0, c1, c2 = chns
        self.inc = _InputParams(in_ch, interm_ch, c0)
        self.downc0 = _DownParams(c0, c1)
        self.downc1 = _DownParams(c1, c2)
        self.upc2 = _UpParams(c2, c1)
        self.upc1 = _UpParams(c1, c0)
        self.outc = _OutParams(c0, out_ch)
        _kaiming_all(self)


class BSVDParams(nn.Module):
    def __init__(self, chns=(32, 64, 128), mid_ch=32, in_ch=4, out_ch=3, interm_ch=30):
        super().__init__()
        self.temp1 = _DenBlockParams(chns, mid_ch, in_ch, interm_ch)
        self.temp2 = _DenBlockParams(chns, out_ch, mid_ch, interm_ch)
        _kaiming_all(self)


def build_bsvd32(seed=0, weight_scale=1.0):
    """State dict of the service's BSVD-32 (bsvd/factory.py:31-35) with the reference constructor's
    random init.  ``weight_scale`` multiplies conv weights (0.5 gives 'trained-like'
    magnitudes whose outputs stay near [0,1])."""
    torch.manual_seed(seed)
    sd = BSVDParams().state_dict()
    if weight_scale != 1.0:
        sd = {k: (v * weight_scale if k.endswith(".weight") else v) for k, v in sd.items()}
    return sd
