"""Micro-benchmark of single convs (ss4k_debug_bench_conv) with pipeline-isolation flags.
Prints one JSON line per configuration: time, TFLOP/s, tile config."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ss4k_b200  # noqa: E402
from ss4k_b200 import _lib as L  # noqa: E402


def bench(eng, cin, cout, h, w, n=1, mode=0, pitch=0, flags=0, iters=20, R=0, slots=0, act=1, beta=0.0, acc=0, grid=0, trace=0, src=0):
    d = L.ConvDesc()
    d.struct_size = ctypes.sizeof(L.ConvDesc)
    d.n, d.h, d.w, d.cin, d.cout, d.mode, d.act = n, h, w, cin, cout, mode, act
    d.alpha, d.beta = 1.0, beta
    d.reserved[1] = src
    d.reserved[2], d.reserved[3], d.reserved[4], d.reserved[5], d.reserved[7] = R, slots, acc, grid, trace
    ms = ctypes.c_float()
    js = ctypes.c_void_p()
    L.check(eng.lib.ss4k_debug_bench_conv(eng.h, ctypes.byref(d), pitch, flags, iters, ctypes.byref(ms), ctypes.byref(js)), eng.h)
    cfg = json.loads(ctypes.string_at(js).decode())
    eng.lib.ss4k_free(js)
    scale = 4 if mode == 1 else (0.25 if mode == 2 else 1)
    flops = 2.0 * cin * cout * 9 * h * w * n * scale
    return {"cin": cin, "cout": cout, "hw": [h, w], "n": n, "mode": mode, "flags": flags, "ms": round(ms.value, 4),
            "tflops": round(flops / ms.value / 1e9, 1), **cfg}


if __name__ == "__main__":
    eng = ss4k_b200.Engine.get(0)
    H, W = 360, 640
    print("# RDB convs, 1 frame 360x640, pipeline-isolation flags (1 no MMA, 2 no TMA, 4 no epilogue math)")
    for cin, cout in [(64, 32), (160, 32), (192, 64), (64, 64)]:
        for flags in (0, 1, 2, 4, 5, 7):
            print(json.dumps(bench(eng, cin, cout, H, W, pitch=192 if cin != 64 or cout == 32 else 0, flags=flags)))
    print("# slab / accumulator ring sweeps")
    for cin, cout in [(64, 32), (160, 32)]:
        for slots in (3, 4, 6, 12):
            print(json.dumps(bench(eng, cin, cout, H, W, pitch=192, slots=slots)))
        for acc in (4, 8, 16):
            print(json.dumps(bench(eng, cin, cout, H, W, pitch=192, acc=acc)))
    print("# batch 4 (bands of ~49 rows)")
    for cin, cout in [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64), (64, 64)]:
        print(json.dumps(bench(eng, cin, cout, H, W, n=4, pitch=192 if cout == 32 or cin == 192 else 0, beta=1.0 if cin == 192 else 0.0)))
    print("# upsampling tail (1 frame)")
    print(json.dumps(bench(eng, 64, 64, 360, 640, mode=1)))
    print(json.dumps(bench(eng, 64, 64, 720, 1280, mode=1)))
    print(json.dumps(bench(eng, 64, 64, 1440, 2560)))
    print(json.dumps(bench(eng, 64, 3, 1440, 2560, act=0)))
    print("# BSVD shapes")
    print(json.dumps(bench(eng, 64, 64, 360, 640, act=2)))
    print(json.dumps(bench(eng, 128, 128, 180, 320, act=2)))
    print(json.dumps(bench(eng, 128, 256, 180, 320, act=0)))
