"""The C-ABI library loads and exports every symbol include/ss4k.h declares; the ctypes table covers them all;
without a GPU the engine refuses to start (no CPU fallback) instead of crashing."""
import ctypes
import os
import re

import ss4k_b200
from ss4k_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ss4k.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss4k_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ss4k.h but not exported by libss4k.so"
        assert n in L.SYMBOLS, f"{n} has no ctypes signature in _lib.SYMBOLS"
    assert lib.ss4k_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = lib.ss4k_create(0, ctypes.byref(h))
    assert rc == -2 and not h.value                       # SS4K_E_NODEVICE
    assert b"no CPU path" in lib.ss4k_last_error(None)
    try:
        ss4k_b200.Engine(0)
        raise AssertionError("Engine() must fail without a GPU")
    except L.Ss4kError:
        pass


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sharkshark-4k_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_plan_cache_lru():
    """Shape-keyed plan cache (still-image server path, image_pipeline.py:54-64: one plan per image size): the least
    recently used plan is closed once max_plans are alive."""
    from ss4k_b200.engine import PlanCache

    closed = []

    class FakePlan:
        def __init__(self, k):
            self.k = k

        def close(self):
            closed.append(self.k)

    c = PlanCache(max_plans=2)
    a = c.get("a", lambda: FakePlan("a"))
    c.get("b", lambda: FakePlan("b"))
    assert c.get("a", lambda: FakePlan("a2")) is a          # hit, and "a" becomes the most recent
    c.get("c", lambda: FakePlan("c"))                       # evicts "b"
    assert closed == ["b"] and "a" in c and "c" in c and len(c) == 2 and c.evictions == 1
    c.get("b", lambda: FakePlan("b"))                       # evicts "a"
    assert closed == ["b", "a"]


def test_mma_issue_loop_keeps_uniform_registers():
    """Code-generation guard for csrc/conv_stream.cu: the MMA-issuing warp's loop must keep its descriptors in uniform
    registers.  A 64-bit division or an out-of-line call anywhere in the kernel makes ptxas fall back to vector
    registers + R2UR moves in front of every tcgen05.mma (measured: 69 -> 270 R2UR, -30 % frames/s)."""
    import shutil
    import subprocess
    import pytest
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "sharkshark-4k_b200", "csrc", "libss4k.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=600).stdout
    blocks = sass.split("Function : ")
    body = [b for b in blocks if b.startswith("_ZN4ss4k21conv3x3_stream_kernelILi32E")]
    assert body, "conv3x3_stream_kernel<32> not found in the library"
    n_mma = body[0].count("UTCHMMA")
    n_r2ur = body[0].count("R2UR")
    assert n_mma >= 24, n_mma
    assert n_r2ur <= 120, f"{n_r2ur} R2UR instructions: the issue loop lost its uniform registers"
