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


# ---------------------------------------------------------------- row-streaming kernel (conv_stream.cu)
from tests.emulate import debug_pack_stream, emulate_stream_conv, emulate_stream_conv_s2  # noqa: E402


@pytest.mark.parametrize("cin,cout,h,w,n,grid,slots", [
    (64, 32, 9, 140, 1, 148, None),     # RDB conv1: bands shorter than the image, two strips
    (96, 32, 21, 130, 2, 5, 4),         # partial second K block, small ring -> many wraps, long bands
    (192, 64, 7, 64, 1, 3, None),       # conv5: Cout split into two 32-wide chunks (weights must fit smem)
    (64, 64, 30, 128, 1, 2, 8),
    (3, 64, 6, 40, 2, 148, None),       # first conv: 16-channel source tensor
    (64, 48, 5, 50, 1, 7, None),        # SRVGG tail (N = 144)
    (64, 3, 12, 300, 1, 16, None),      # conv_last: Cout padded to 16
    (128, 256, 4, 70, 3, 11, None),     # BSVD upc2: four 64-wide chunks, weights reloaded per chunk
    (30, 32, 40, 20, 1, 4, 3),          # minimal ring of three slots
])
def test_stream_schedule(lib, cin, cout, h, w, n, grid, slots):
    wt = _exact_weights(cout, cin, 11)
    x = torch.randint(-8, 9, (n, cin, h, w), generator=torch.Generator().manual_seed(12)).float() / 8
    pitch = (cin + 15) // 16 * 16
    bias = torch.randint(-64, 65, (cout,), generator=torch.Generator().manual_seed(15)).float() / 32   # exact in fp16
    meta, packed = debug_pack_stream(lib, L, wt, bias=bias)
    got = emulate_stream_conv(meta, packed, _nhwc(x, pitch), grid, acc_slots=slots)
    want = F.conv2d(x.double(), wt.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert torch.equal(got[..., :cout], want)
    assert torch.count_nonzero(got[..., cout:]) == 0
    assert meta["nkb"] == (cin + 63) // 64


def test_stream_channel_offset_and_permutation(lib):
    """Input at a channel offset of a wider slab (tensor-map base shift) and the PixelShuffle(2) row order."""
    wt = _exact_weights(64, 64, 13)
    slab = torch.randint(-8, 9, (1, 6, 33, 192), generator=torch.Generator().manual_seed(14)).float() / 8
    meta, packed = debug_pack_stream(lib, L, wt, in_pitch=192, in_coff=64, wperm=1)
    got = emulate_stream_conv(meta, packed, slab, 9, in_coff=64)
    want = F.conv2d(slab[..., 64:128].permute(0, 3, 1, 2).double(), wt.double(), padding=1).permute(0, 2, 3, 1)
    perm = [(r % 16) * 4 + r // 16 for r in range(64)]   # packed row (a*2+b)*16 + c  <-  channel c*4 + a*2 + b
    assert torch.equal(got, want[..., perm])


def test_stream_alpha_fold_and_bias(lib):
    """alpha is folded into weights and bias for none / PReLU activations (not for ReLU6); the bias is carried
    in fp32 (the accumulators' initial value), so it need not be representable in fp16."""
    wt = _exact_weights(32, 64, 21)
    bias = torch.full((32,), 0.1234567)
    meta, packed = debug_pack_stream(lib, L, wt, bias=bias, alpha=0.25, act=1)
    assert torch.equal(packed[0, 0, 1, 1, :, :], (0.25 * wt[:, :, 1, 1]))
    bf = torch.tensor(meta["bias_f"])
    assert bf.shape == (32,) and abs(bf.double() - 0.25 * 0.1234567).max().item() < 2e-8
    meta6, packed6 = debug_pack_stream(lib, L, wt, bias=bias, alpha=0.25, act=2)
    assert torch.equal(packed6[0, 0, 1, 1, :, :], wt[:, :, 1, 1])


@pytest.mark.parametrize("cin,cout,h,w,n,grid,slots", [
    (32, 64, 16, 40, 1, 148, None),     # BSVD downc0: one K block = [even pixel 32 ch | odd pixel 32 ch]
    (64, 128, 12, 300, 2, 7, None),     # BSVD downc1: two K blocks, two 64-wide chunks, two strips
    (32, 64, 44, 24, 1, 3, 4),          # long bands, small ring -> wraps inside the two-row windows
    (64, 64, 8, 520, 1, 5, 3),          # minimal ring, three strips
])
def test_stream_schedule_stride2(lib, cin, cout, h, w, n, grid, slots):
    """Stride-2 convs on the row-streaming kernel (pixel-pair view, [ky2 | ky0 | ky1] weight blocks)."""
    wt = _exact_weights(cout, cin, 31)
    x = torch.randint(-8, 9, (n, cin, h, w), generator=torch.Generator().manual_seed(32)).float() / 8
    bias = torch.randint(-64, 65, (cout,), generator=torch.Generator().manual_seed(33)).float() / 32
    meta, packed = debug_pack_stream(lib, L, wt, bias=bias, mode=2, in_pitch=cin)
    assert meta["stride2"] == 1 and meta["nkb"] == 2 * cin // 64
    got = emulate_stream_conv_s2(meta, packed, _nhwc(x, cin), grid, acc_slots=slots)
    want = F.conv2d(x.double(), wt.double(), bias.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert torch.equal(got[..., :cout], want)


def test_stream_stride2_split_blocks(lib):
    """fp16 hi/lo split operands on the stride-2 path: three K blocks per merged 64-channel block."""
    wt = _exact_weights(64, 32, 41)
    meta, packed = debug_pack_stream(lib, L, wt, mode=2, in_pitch=32, act_mode=L.ACT_F16_SPLIT)
    assert meta["nkb"] == 3 and meta["src_kb"] == [0, 0, 0] and meta["ksm"] == [[12, 15]] * 3
    assert meta["nwt"] == 2 and meta["wt"] == [0, 1, 0]      # A_hi*W_hi and A_lo*W_hi share ONE weight tile group
    assert torch.equal(packed[0, 0], packed[0, 2])
    assert torch.count_nonzero(packed[0, 1]) == 0             # exactly representable weights: the low halves are zero
