"""GPU parity tests of the tcgen05 convolution through the C ABI (operator-level entry ss4k_conv3x3)
against the CPU oracle meaning of the same op (torch fp32 conv on the 16-bit-rounded operands)."""
import pytest
import torch
import torch.nn.functional as F

from ss4k_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _ref(x, w, b, mode, act, slope, alpha, beta, res, ps, dt=torch.float16):
    xq = x.to(dt).double()
    wq = w.to(dt).double()
    if mode == L.MODE_UP2:
        # the kernel pre-sums the taps that collapse onto one low-res pixel (fp32) and rounds once
        v = F.conv2d(F.interpolate(xq, scale_factor=2, mode="nearest"), w.double(), b.double() if b is not None else None, padding=1)
    elif mode == L.MODE_S2:
        v = F.conv2d(xq, wq, b.double() if b is not None else None, stride=2, padding=1)
    else:
        v = F.conv2d(xq, wq, b.double() if b is not None else None, padding=1)
    if act == 1:
        v = torch.where(v >= 0, v, v * slope.double().view(1, -1, 1, 1))
    elif act == 2:
        v = v.clamp(0, 6)
    v = v * alpha
    if ps:
        v = F.pixel_shuffle(v, ps)
    if res is not None:
        v = v + beta * res.to(dt).double()
    return v.float()


CASES = [
    # cin, cout, h, w, n, mode, act, residual, ps, direct
    (64, 64, 7, 200, 1, 0, 0, False, 0, True),      # the self-probe shape
    (64, 64, 7, 200, 1, 0, 1, False, 0, False),     # NHWC store path + PReLU
    (3, 64, 9, 130, 2, 0, 1, False, 0, False),      # first conv, channel-padded input, batch 2
    (12, 64, 5, 64, 1, 0, 0, False, 0, False),
    (96, 32, 6, 140, 1, 0, 1, False, 0, False),     # RDB growth conv (partial second K block)
    (160, 32, 6, 140, 1, 0, 1, False, 0, False),
    (192, 64, 6, 140, 1, 0, 0, True, 0, False),     # RDB conv5 + scaled residual, weights double-buffered
    (64, 3, 10, 260, 1, 0, 0, False, 0, True),      # conv_last: N=16 accumulator, float NCHW store
    (64, 48, 6, 70, 1, 0, 0, False, 4, True),       # SRVGG tail: PixelShuffle(4) float store
    (64, 64, 5, 130, 1, 1, 1, False, 0, False),     # nearest-x2 fused (4 phases)
    (32, 64, 8, 264, 1, 2, 2, False, 0, False),     # stride 2 + ReLU6 (BSVD downc0)
    (64, 128, 8, 264, 1, 2, 2, False, 0, False),    # stride 2, two N chunks (BSVD downc1)
    (128, 256, 4, 140, 1, 0, 0, True, 2, False),    # BSVD upc2: 4 N chunks + PixelShuffle(2) + skip add
    (30, 32, 6, 140, 1, 0, 2, False, 0, False),     # BSVD inc second conv (30 -> 32)
    (64, 64, 200, 640, 1, 0, 1, False, 0, False),   # 250 tiles > 148 SMs: persistent loop, TMEM double buffer
]


@pytest.mark.parametrize("cin,cout,h,w,n,mode,act,use_res,ps,direct", CASES)
def test_conv_parity(engine, cin, cout, h, w, n, mode, act, use_res, ps, direct):
    g = torch.Generator().manual_seed(cin * 1000 + cout + h)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    slope = torch.rand(cout, generator=g) * 0.5 if act == 1 else None
    alpha, beta = (0.2, 1.0) if use_res else (1.0, 0.0)
    oh, ow = (2 * h, 2 * w) if mode == 1 else ((h // 2, w // 2) if mode == 2 else (h, w))
    oc = cout // (ps * ps) if ps else cout
    res = torch.randn(n, oc, oh * (ps or 1), ow * (ps or 1), generator=g) if use_res else None
    want = _ref(x, wt, b, mode, act, slope, alpha, beta, res, ps)
    y = engine.conv3x3(x.cuda(), wt.cuda(), b.cuda(), slope.cuda() if slope is not None else None,
                       res.cuda() if res is not None else None, mode=mode, act=act, pixel_shuffle=ps,
                       alpha=alpha, beta=beta, direct_f32=direct)
    torch.cuda.synchronize()
    got = y.cpu()
    assert got.shape == want.shape
    # fp32 accumulation order differs; NHWC paths also round the result to fp16 once
    tol = 2e-3 if direct else 4e-3
    err = (got - want).abs().max().item()
    assert err <= tol * max(1.0, want.abs().max().item()), (err, want.abs().max().item())


def test_conv_bf16(engine):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 6, 150, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    b = torch.randn(64, generator=g) * 0.1
    want = _ref(x, wt, b, 0, 0, None, 1.0, 0.0, None, 0, dt=torch.bfloat16)
    got = engine.conv3x3(x.cuda(), wt.cuda(), b.cuda(), act_mode=L.ACT_BF16, direct_f32=True).cpu()
    assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


def test_conv_split3(engine):
    """fp16 hi/lo split operands: error vs the exact fp32 conv drops by > 50x (BSVD precision mode)."""
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 64, 6, 150, generator=g) * 3
    wt = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    b = torch.randn(64, generator=g) * 0.1
    exact = F.conv2d(x.double(), wt.double(), b.double(), padding=1).float()
    single = engine.conv3x3(x.cuda(), wt.cuda(), b.cuda(), direct_f32=True).cpu()
    split = engine.conv3x3(x.cuda(), wt.cuda(), b.cuda(), act_mode=L.ACT_F16_SPLIT, direct_f32=True).cpu()
    e1 = (single - exact).abs().max().item()
    e3 = (split - exact).abs().max().item()
    assert e3 < 1e-4 and e3 < e1 / 20, (e1, e3)


def test_engine_reports_mode_and_launches(engine):
    assert engine.desc_mode in (0, 1, 2)
    before = engine.launch_count
    x = torch.randn(1, 64, 4, 130).cuda()
    engine.conv3x3(x, torch.randn(64, 64, 3, 3).cuda() * 0.05, direct_f32=True)
    assert engine.launch_count >= before + 2


@pytest.mark.parametrize("cin,cout,use_res,act_mode", [
    (64, 128, True, L.ACT_F16), (64, 128, True, L.ACT_F16_SPLIT), (64, 128, False, L.ACT_F16),
    (128, 256, True, L.ACT_F16_SPLIT),          # a 32-wide chunk is half of a 64-channel sub-pixel phase
])
def test_conv_pixelshuffle2_skip_tma_store(engine, cin, cout, use_res, act_mode, monkeypatch):
    """BSVD up-convs (bsvd/model.py:290-323): conv + PixelShuffle(2) + skip add leaving through the staging tile and 5-D
    TMA stores (one sub-pixel phase per chunk; twin hi / lo tiles in split precision).  Ragged width, two frames; against
    the fp64 reference and against the per-thread store path (SS4K_NO_PS2_FAST=1: that plan may pick a 64-wide chunk, whose
    bias enters through an MMA as hi + lo halves instead of the fp32 accumulator init -- equal to one fp16 rounding)."""
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(2, cin, 6, 140, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(2, cout // 4, 12, 280, generator=g) if use_res else None
    want = _ref(x, wt, b, 0, 0, None, 1.0, 1.0, res, 2)

    def run():
        return engine.conv3x3(x.cuda(), wt.cuda(), b.cuda(), None, res.cuda() if use_res else None, act=0, pixel_shuffle=2,
                              alpha=1.0, beta=1.0, act_mode=act_mode).cpu()

    fast = run()
    monkeypatch.setenv("SS4K_NO_PS2_FAST", "1")
    slow = run()
    assert fast.shape == want.shape
    assert (fast - slow).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    assert (fast - want).abs().max().item() <= 4e-3 * max(1.0, want.abs().max().item())
