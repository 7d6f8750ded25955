"""Weight-loading code on the product path, on the CPU (no engine is created):
  * realesrgan.dni / load_checkpoint / build_model's DNI branches (reference: realesrgan/factory.py:152-170,
    RealESRGANer.dni + load)
  * bsvd.load_checkpoint's key remap against the reference's OWN BSVD.load (bsvd/model.py:487-499 with the
    DownBlock / UpBlock / MemCvBlock renames at :167-169,276-279,304-306) where /root/reference exists, and against
    the oracle's key list everywhere."""
import os

import pytest
import torch

from ss4k_b200 import bsvd as native_bsvd
from ss4k_b200 import realesrgan
from oracle import bsvd as obsvd
from oracle import reference_import as ri
from oracle import srvgg


def _sd(seed, num_conv=2):
    torch.manual_seed(seed)
    return srvgg.SRVGGNetCompact(3, 3, 64, num_conv, 4).state_dict()


def test_dni_is_the_keywise_blend():
    a, b = _sd(1), _sd(2)
    for s in (0.75, 0.2, 0.5, 1.0, 0.0):
        w = realesrgan.dni(a, b, [s, 1 - s])
        assert list(w) == list(a)
        for k in a:
            assert torch.equal(w[k], s * a[k] + (1 - s) * b[k])


def test_load_checkpoint_prefers_params_ema(tmp_path):
    a, b = _sd(1), _sd(2)
    p = tmp_path / "x.pth"
    torch.save({"params": a, "params_ema": b}, p)
    got = realesrgan.load_checkpoint(str(p))
    assert all(torch.equal(got[k], b[k]) for k in b)
    torch.save({"params": a}, p)
    got = realesrgan.load_checkpoint(str(p))
    assert all(torch.equal(got[k], a[k]) for k in a)


class _Capture:
    """stands in for the engine-backed module: records the weights build_model hands to it"""
    last = None

    def __init__(self, state_dict, **kw):
        _Capture.last = (state_dict, kw)

    def eval(self):
        return self


@pytest.fixture
def captured(monkeypatch):
    monkeypatch.setattr(realesrgan, "NativeSRVGG", _Capture)
    monkeypatch.setattr(realesrgan, "NativeRRDBNet", _Capture)
    return _Capture


def test_build_model_blends_general_and_wdn(tmp_path, captured):
    """the reference's live callers use denoise_rate 0.75 / 0.2 with realesr-general-x4v3: both weight sets are blended,
    from a (general, wdn) state-dict pair, from a model_path pair and from the wdn file next to model_path"""
    a, b = _sd(1, 32), _sd(2, 32)
    for s in (0.75, 0.2):
        want = {k: s * a[k] + (1 - s) * b[k] for k in a}
        realesrgan.build_model(denoise_rate=s, state_dict=(a, b))
        got, kw = captured.last
        assert all(torch.equal(got[k], want[k]) for k in want) and kw["num_conv"] == 32
        pa, pb = tmp_path / "realesr-general-x4v3.pth", tmp_path / "realesr-general-wdn-x4v3.pth"
        torch.save({"params": a}, pa)
        torch.save({"params": b}, pb)
        args = realesrgan.ArgsData()
        args.model_path = str(pa)
        realesrgan.build_model(denoise_rate=s, args=args)
        got, _ = captured.last
        assert all(torch.equal(got[k], want[k]) for k in want)
        args.model_path = [str(pa), str(pb)]
        realesrgan.build_model(denoise_rate=s, args=args)
        got, _ = captured.last
        assert all(torch.equal(got[k], want[k]) for k in want)
    # denoise_strength == 1: the general weights alone (factory.py:154)
    realesrgan.build_model(denoise_rate=1, state_dict=a)
    got, _ = captured.last
    assert all(torch.equal(got[k], a[k]) for k in a)


def test_build_model_refuses_a_silent_unblended_net(tmp_path, captured):
    a = _sd(1, 32)
    with pytest.raises(ValueError, match="wdn"):
        realesrgan.build_model(denoise_rate=0.75, state_dict=a)
    pa = tmp_path / "realesr-general-x4v3.pth"
    torch.save({"params": a}, pa)
    args = realesrgan.ArgsData()
    args.model_path = str(pa)
    with pytest.raises(FileNotFoundError, match="wdn"):
        realesrgan.build_model(denoise_rate=0.2, args=args)
    # other models never blend
    args2 = realesrgan.ArgsData()
    args2.model_name = 'realesr-animevideov3'
    realesrgan.build_model(denoise_rate=0.75, args=args2, state_dict=_sd(3, 16))
    assert captured.last[1]["num_conv"] == 16


# ---------------------------------------------------------------------------------------------- BSVD checkpoint
def _to_checkpoint_keys(model_sd, prefix):
    """inverse of the reference's load-time renames: model key -> training-time checkpoint key"""
    out = {}
    for k, v in model_sd.items():
        for i, t in ((0, "temp1."), (1, "temp2.")):
            if not k.startswith(t):
                continue
            blk, rest = k[len(t):].split(".", 1)
            if blk in ("downc0", "downc1") and rest.startswith("memconv."):
                rest = "convblock.3." + rest[len("memconv."):].replace("op.conv.", "net.")
            elif blk in ("upc2", "upc1"):
                if rest.startswith("memconv."):
                    rest = "convblock.0." + rest[len("memconv."):].replace("op.conv.", "net.")
                elif rest.startswith("convblock.0."):
                    rest = "convblock.1." + rest[len("convblock.0."):]
            out[f"{prefix}nets_list.{i}.{blk}.{rest}"] = v
    return out


@pytest.mark.parametrize("prefix", ["base_model.", "module.base_model."])
def test_bsvd_load_checkpoint_key_remap(tmp_path, prefix):
    want = obsvd.build_bsvd32(3)
    p = tmp_path / "bsvd-32.pth"
    torch.save({"params": _to_checkpoint_keys(want, prefix)}, p)
    got = native_bsvd.load_checkpoint(str(p))
    assert sorted(got) == sorted(want)
    assert all(torch.equal(got[k], want[k]) for k in want)


@pytest.mark.skipif(not ri.available(), reason="/root/reference not present")
@pytest.mark.parametrize("prefix", ["base_model.", "module.base_model."])
def test_bsvd_load_checkpoint_matches_reference_load(tmp_path, prefix):
    """the reference's own BSVD.load on the same synthetic checkpoint ends up with exactly the tensors our remap
    returns under the reference model's own state-dict key names"""
    m = ri.load_bsvd_model()
    torch.manual_seed(11)
    with ri.cpu_shims():
        src = m.BSVD(chns=[32, 64, 128], mid_ch=32, shift_input=False, norm='none', interm_ch=30, act='relu6', pretrain_ckpt=None)
    ck = _to_checkpoint_keys(src.state_dict(), prefix)
    assert len(ck) == len(src.state_dict())
    p = tmp_path / "bsvd-32.pth"
    torch.save({"params": ck}, p)
    torch.manual_seed(12)   # a differently initialised model: every tensor must come from the checkpoint
    with ri.cpu_shims():
        ref = m.BSVD(chns=[32, 64, 128], mid_ch=32, shift_input=False, norm='none', interm_ch=30, act='relu6', pretrain_ckpt=str(p))
    ref_sd = ref.state_dict()
    got = native_bsvd.load_checkpoint(str(p))
    assert sorted(got) == sorted(ref_sd)
    for k, v in ref_sd.items():
        assert torch.equal(got[k], v), k
        assert torch.equal(v, src.state_dict()[k]), k
