"""CPU checks of the weight packer and MMA schedule tables (csrc/engine.cu pack_weights,
configure_tiles) by emulating the kernel's tap/K-block schedule from ss4k_debug_pack output."""
import pytest
import torch
import torch.nn.functional as F

import ss4k_b200
from ss4k_b200 import _lib as L
from tests.emulate import debug_pack, emulate_packed_conv


def _exact_weights(cout, cin, seed):
    g = torch.Generator().manual_seed(seed)
    # multiples of 1/64 in [-1,1]: exactly representable in fp16, and so are sums of up to 9 of them
    return torch.randint(-64, 65, (cout, cin, 3, 3), generator=g).float() / 64.0


def _nhwc(x, pitch):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, pitch)
    out[..., :c] = x.permute(0, 2, 3, 1)
    return out


@pytest.mark.parametrize("cin,cout", [(64, 64), (3, 64), (12, 64), (96, 32), (192, 64), (64, 48), (64, 3), (30, 32), (128, 128)])
def test_conv3_schedule(lib, cin, cout):
    w = _exact_weights(cout, cin, 1)
    x = torch.randint(-8, 9, (1, cin, 6, 9), generator=torch.Generator().manual_seed(2)).float() / 8
    meta, packed = debug_pack(lib, L, w, mode=0)
    got = emulate_packed_conv(meta, packed, _nhwc(x, (cin + 15) // 16 * 16), 0)
    want = F.conv2d(x.double(), w.double(), padding=1).permute(0, 2, 3, 1)
    assert torch.equal(got[..., :cout], want)
    assert torch.count_nonzero(got[..., cout:]) == 0
    assert meta["nkb"] == (cin + 63) // 64 and meta["ntaps"] == 9


def test_conv3_channel_offset(lib):
    """RDB growth conv reading channels [0, 96) of a 192-pitch slab, like body.k.rdbN.conv2."""
    w = _exact_weights(32, 96, 3)
    slab = torch.randint(-8, 9, (1, 5, 7, 192), generator=torch.Generator().manual_seed(4)).float() / 8
    meta, packed = debug_pack(lib, L, w, mode=0, in_pitch=192, in_coff=0)
    got = emulate_packed_conv(meta, packed, slab, 0)
    want = F.conv2d(slab[..., :96].permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1)
    assert torch.equal(got[..., :32], want)
    # second k-block holds only 32 real channels: k-steps 2,3 must be masked off (stale slab data there)
    assert all(m == 0b0011 for m in meta["mask"][1])


def test_up2_schedule(lib):
    """nearest-x2 upsample + 3x3 conv == 4 output phases x 4 pre-summed taps on the low-res image."""
    w = _exact_weights(64, 64, 5)
    x = torch.randint(-8, 9, (1, 64, 5, 6), generator=torch.Generator().manual_seed(6)).float() / 8
    meta, packed = debug_pack(lib, L, w, mode=1)
    assert meta["ntaps"] == 16 and meta["nsub"] == 4
    got = emulate_packed_conv(meta, packed, _nhwc(x, 64), 1)
    want = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), w.double(), padding=1).permute(0, 2, 3, 1)
    assert torch.equal(got, want)


@pytest.mark.parametrize("cin,cout", [(32, 64), (64, 128)])
def test_stride2_schedule(lib, cin, cout):
    """3x3 stride-2 conv read through the (2C, W/2, 2, H/2) view with k-step masks (BSVD DownBlock)."""
    w = _exact_weights(cout, cin, 7)
    x = torch.randint(-8, 9, (1, cin, 8, 12), generator=torch.Generator().manual_seed(8)).float() / 8
    meta, packed = debug_pack(lib, L, w, mode=2, in_pitch=cin)
    got = emulate_packed_conv(meta, packed, _nhwc(x, cin), 2)
    want = F.conv2d(x.double(), w.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert torch.equal(got[..., :cout], want)
    # no wasted k-steps: exactly 9 taps x cin/16 k-steps are issued
    issued = sum(bin(m).count("1") for row in meta["mask"] for m in row)
    assert issued == 9 * cin // 16


def test_pixel_shuffle_permutation(lib):
    """wperm=1: packed output rows ordered (a,b,c) so PixelShuffle(2) becomes a plain NHWC store."""
    w = _exact_weights(128, 64, 9)
    b = torch.arange(128).float()
    meta, packed = debug_pack(lib, L, w, bias=b, mode=0, wperm=1)
    x = torch.randint(-8, 9, (1, 64, 4, 5), generator=torch.Generator().manual_seed(10)).float() / 8
    got = emulate_packed_conv(meta, packed, _nhwc(x, 64), 0) + torch.tensor(meta["bias"]).double()
    conv = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    want = F.pixel_shuffle(conv, 2)                       # [1, 32, 8, 10]
    # packed column (a*2+b)*32 + c at (y, x)  ==  shuffled[c, 2y+a, 2x+b]
    for a in range(2):
        for bb in range(2):
            blk = got[..., (a * 2 + bb) * 32:(a * 2 + bb + 1) * 32]
            assert torch.equal(blk, want[:, :, a::2, bb::2].permute(0, 2, 3, 1))


def test_split_operands(lib):
    """fp16 hi/lo split: 3 K blocks per 64 channels (A_hi*W_hi + A_hi*W_lo + A_lo*W_hi)."""
    g = torch.Generator().manual_seed(11)
    w = torch.randn(32, 64, 3, 3, generator=g) * 0.1
    x = torch.randn(1, 64, 5, 6, generator=g)
    meta, packed = debug_pack(lib, L, w, mode=0, act_mode=L.ACT_F16_SPLIT)
    assert meta["nkb"] == 3 and [k[0] for k in meta["kb"]] == [0, 0, 1]
    xh = x.half().float()
    xl = (x - xh).half().float()
    got = emulate_packed_conv(meta, packed, _nhwc(xh, 64), 0, x_lo=_nhwc(xl, 64))
    want = F.conv2d(x.double(), w.double(), padding=1).permute(0, 2, 3, 1)
    single = F.conv2d(xh.double(), w.half().double(), padding=1).permute(0, 2, 3, 1)
    err_split = (got[..., :32] - want).abs().max().item()
    err_single = (single - want).abs().max().item()
    assert err_split < 2e-5 and err_split < err_single / 50


@pytest.mark.parametrize("desc_mode", [0, 2])
def test_tile_config_budget(lib, desc_mode):
    """Shared-memory / TMEM budgets of every conv shape on the hot path."""
    shapes = [(64, 64, 0, 360, 640), (192, 64, 0, 360, 640), (160, 32, 0, 360, 640), (64, 64, 1, 720, 1280),
              (64, 3, 0, 1440, 2560), (64, 48, 0, 180, 320), (16, 64, 0, 180, 320), (128, 256, 0, 180, 320),
              (64, 128, 2, 360, 640)]
    for cin, cout, mode, h, w in shapes:
        wt = torch.zeros(cout, cin, 3, 3)
        meta, _ = debug_pack(lib, L, wt, mode=mode, h=h, wd=w, desc_mode=desc_mode,
                             in_pitch=cin if mode == 2 else None)
        smem = meta["w_slots"] * meta["w_slot_bytes"] + meta["a_slots"] * meta["a_slot_bytes"] + 1024 + 512
        assert smem <= 232448, (cin, cout, mode, meta)
        assert meta["a_slots"] >= 2
        assert meta["R"] * meta["nsub"] * meta["acc_stride"] <= 256
        assert meta["n_cta"] % 16 == 0 and meta["n_cta"] <= 64
