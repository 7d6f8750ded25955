"""CPU checks of the network lowering (csrc/program.cpp): the layer program emitted by ss4k_plan_dry is
interpreted with torch fp32 ops and compared with the oracle on the same seeded weights / inputs."""
import pytest
import torch

import ss4k_b200
from ss4k_b200 import _lib as L
from oracle import rrdbnet, srvgg
from tests.emulate import run_program


@pytest.mark.parametrize("num_conv,upscale", [(16, 4), (32, 4), (4, 2)])
def test_srvgg_program(lib, num_conv, upscale):
    torch.manual_seed(0)
    net = srvgg.SRVGGNetCompact(3, 3, 64, num_conv, upscale).eval()
    x = torch.rand(2, 3, 20, 28)
    prog = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_SRVGG, 2, 20, 28, scale=upscale, depth=num_conv))
    assert len(prog["steps"]) == 1 + num_conv + 2
    with torch.no_grad():
        want = net(x)
        got = run_program(prog, net.state_dict(), x)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 1e-4
    # algorithmic FLOPs: 2*9*HW*(3*64 + num_conv*64*64 + 64*3*s*s) per frame
    flops = 2 * 9 * 20 * 28 * 2 * (3 * 64 + num_conv * 64 * 64 + 64 * 3 * upscale * upscale)
    assert abs(prog["flops"] - flops) / flops < 1e-9


@pytest.mark.parametrize("scale,blocks", [(2, 2), (4, 1)])
def test_rrdb_program(lib, scale, blocks):
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, scale, 64, blocks, 32).eval()
    # upstream init makes RDB convs tiny; also exercise non-zero biases
    for p in net.parameters():
        if p.ndim == 1:
            torch.nn.init.uniform_(p, -0.1, 0.1)
    x = torch.rand(1, 3, 16, 24)
    prog = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_RRDB, 1, 16, 24, scale=scale, depth=blocks))
    with torch.no_grad():
        want = net(x)
        got = run_program(prog, net.state_dict(), x)
    assert got.shape == want.shape == (1, 3, 16 * scale, 24 * scale)
    assert (got - want).abs().max().item() < 1e-4


def test_rrdb_flops_match_baseline(lib):
    """RRDBNet-23 x2 @ 1280x720 = 8.263 TFLOP, x4 @ 1920x1080 = 74.346 TFLOP."""
    p2 = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_RRDB, 1, 720, 1280, scale=2, depth=23))
    assert abs(p2["flops"] / 1e12 - 8.263) < 0.005
    p4 = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_RRDB, 1, 1080, 1920, scale=4, depth=23))
    assert abs(p4["flops"] / 1e12 - 74.346) < 0.02
    ps = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_SRVGG, 1, 180, 320, scale=4, depth=32))
    assert abs(ps["flops"] / 1e12 - 0.139) < 0.001


def test_rrdb_fp16_storage_error_budget(lib):
    """Emulated fp16 activation storage through the lowered program stays inside the parity gate
    (PSNR >= 50 dB, max abs <= 2/255) for upstream-init RRDBNet -- SURVEY.md section 7 H2."""
    torch.manual_seed(0)
    net = rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval()
    x = torch.rand(1, 3, 32, 32)
    prog = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_RRDB, 1, 32, 32, scale=2, depth=23))
    sd = {k: (v.half().float() if v.ndim == 4 else v) for k, v in net.state_dict().items()}
    with torch.no_grad():
        want = net(x).clamp(0, 1)
        got = run_program(prog, sd, x, quant=lambda t: t.half().float()).clamp(0, 1)
    mse = torch.mean((got - want) ** 2).item()
    psnr = -10 * torch.log10(torch.tensor(mse)).item()
    assert psnr >= 50 and (got - want).abs().max().item() * 255 <= 2.0


@pytest.mark.parametrize("frames", [1, 5])
def test_bsvd_program(lib, frames):
    """BSVD clip lowering (temporal-shift scatter stores, PixelShuffle + skip adds, none_minus as negated
    weights + masked residual) against the oracle clip function, i.e. the reference BSVD.forward."""
    from oracle import bsvd
    sd = bsvd.build_bsvd32(0)
    x = torch.rand(1, frames, 4, 16, 24, generator=torch.Generator().manual_seed(1234))
    x[:, :, 3] = 0.075
    prog = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_BSVD, frames, 16, 24))
    assert sum(1 for s in prog["steps"] if s["kind"] == "conv") == 32
    want = bsvd.bsvd_forward(sd, x)[0]
    got = run_program(prog, sd, x)
    assert got.shape == want.shape == (frames, 3, 16, 24)
    assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())
    # algorithmic FLOPs: 272.0 GMAC per 1280x720 frame
    p720 = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_BSVD, 1, 720, 1280))
    assert abs(p720["flops"] / 2e9 - 271.99) < 0.05


@pytest.mark.parametrize("T,lo,hi", [(40, 16, 24), (30, 0, 14), (30, 14, 30), (9, 3, 6)])
def test_bsvd_program_owned_frames_with_halo(lib, T, lo, hi):
    """A chunk of a sharded stream: frames [lo, hi) are owned, the rest is temporal halo (sharding.bsvd_chunks).  Every
    layer then runs only on the frames the owned outputs depend on (owned range widened by the number of shift convs
    between the layer and the output): the result must equal the owned frames of the whole-chunk program exactly, and
    with a full 16-frame halo on both sides about half of the halo work must be gone."""
    from oracle import bsvd
    sd = bsvd.build_bsvd32(0)
    x = torch.rand(1, T, 4, 8, 16, generator=torch.Generator().manual_seed(7))
    x[:, :, 3] = 0.075
    full_cfg = ss4k_b200.make_cfg(0, L.ARCH_BSVD, T, 8, 16)
    full = ss4k_b200.plan_dry(full_cfg)
    cfg = ss4k_b200.make_cfg(0, L.ARCH_BSVD, T, 8, 16)
    cfg.reserved[2], cfg.reserved[3] = lo, hi
    prog = ss4k_b200.plan_dry(cfg)
    assert prog["out_n"] == hi - lo and prog["in_n"] == T
    want = run_program(full, sd, x)[lo:hi]
    got = run_program(prog, sd, x)
    assert got.shape == want.shape
    assert torch.equal(got, want)
    convs = [s for s in prog["steps"] if s["kind"] == "conv"]
    assert convs[-1]["n0"] == lo and convs[-1]["n"] == hi - lo               # the last conv: owned frames only
    assert convs[0]["n0"] == max(0, lo - 16) and convs[0]["n0"] + convs[0]["n"] == min(T, hi + 16)   # 16 shift convs to the output
    assert all(a["n"] >= b["n"] for a, b in zip(convs, convs[1:]))           # the ranges only shrink towards the output
    if (T, lo, hi) == (40, 16, 24):
        halo_full = full["flops"] * (T - (hi - lo)) / T
        halo_now = prog["flops"] - full["flops"] * (hi - lo) / T
        assert 0.4 < halo_now / halo_full < 0.6


def test_rrdb_memory_management_is_sound(lib):
    """L2 management of the dense block checked against a liveness analysis of the program:
    * a conv that DISCARDS 128-byte lines of a slab (discard_buf / discard_mask: line l = channels [64 l, 64 l + 64))
      must come after the last step that reads those channels and before the next step that writes them;
    * old_cin (channels a conv may load before the dependency wait) must only cover channels whose last writer is at
      least two steps back."""
    prog = ss4k_b200.plan_dry(ss4k_b200.make_cfg(0, L.ARCH_RRDB, 1, 32, 48, scale=2, depth=3))
    steps = prog["steps"]
    convs = [(i, s) for i, s in enumerate(steps) if s["kind"] == "conv"]

    def reads(s):   # (buf, c0, c1) ranges read by a conv: input channels and residuals
        r = [(s["in_buf"], s["in_coff"], s["in_coff"] + s["cin"])]
        if s["res1_buf"] >= 0:
            r.append((s["res1_buf"], s["res1_coff"], s["res1_coff"] + s["cout"]))
        if s["res2_buf"] >= 0:
            r.append((s["res2_buf"], s["res2_coff"], s["res2_coff"] + s["cout"]))
        return r

    def writes(s):
        return (s["out_buf"], s["out_coff"], s["out_coff"] + s["cout"])

    def overlap(a, b):
        return a[0] == b[0] and a[1] < b[2] and b[1] < a[2]

    n_discards = 0
    for i, s in convs:
        if s["discard_buf"] >= 0 and s["discard_mask"]:
            n_discards += 1
            for l in range(3):
                if not (s["discard_mask"] >> l) & 1:
                    continue
                dead = (s["discard_buf"], 64 * l, 64 * l + 64)
                # every later step that reads these channels must be preceded by a writer that comes after the discard
                rewritten = set()
                for j, t in convs:
                    if j < i:
                        continue
                    for r in reads(t):
                        if overlap(r, dead):
                            lo, hi = max(r[1], dead[1]), min(r[2], dead[2])
                            assert all(c in rewritten for c in range(lo, hi)), (s["name"], "discards", dead, "read by", t["name"])
                    w = writes(t)
                    if overlap(w, dead):
                        rewritten.update(range(max(w[1], dead[1]), min(w[2], dead[2])))
        if s["old_cin"] > 0:
            # last writer of every channel in [0, old_cin) of the input tensor is at least two conv steps back
            pos = [k for k, (j, _) in enumerate(convs) if j == i][0]
            prev = convs[pos - 1][1]
            w = writes(prev)
            assert not overlap(w, (s["in_buf"], s["in_coff"], s["in_coff"] + s["old_cin"])), (s["name"], prev["name"])
            assert s["old_cin"] <= s["cin"]
    assert n_discards == 3 * 3 - 1          # every RDB but the first drops its predecessor's slab
    hints = {(s["l2_in"], s["l2_out"]) for _, s in convs if s["name"].startswith("body.")}
    assert hints == {(1, 1), (2, 1)}
