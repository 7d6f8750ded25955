"""Diagnostic for the tcgen05 conv kernel (run on a B200 via gpurun).  For every shared-memory
descriptor mode (own subprocess: a trap poisons the CUDA context) it runs single-tap convolutions and
reports, per tap, whether the output equals the correspondingly shifted input -- which pins down
descriptor / swizzle / shift mistakes from one run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mode):
    import torch
    import ss4k_b200
    from ss4k_b200 import _lib as L
    import torch.nn.functional as F
    eng = ss4k_b200.Engine.get(0)
    res = {"mode": mode, "desc_mode": eng.desc_mode, "taps": {}}
    g = torch.Generator().manual_seed(0)
    C, H, W = 64, 6, 200
    x = torch.randn(1, C, H, W, generator=g)
    xq = x.half().float()
    for ky in range(3):
        for kx in range(3):
            w = torch.zeros(C, C, 3, 3)
            for c in range(C):
                w[c, (c * 7 + 3) % C, ky, kx] = 1.0          # a channel permutation at one tap
            y = eng.conv3x3(x.cuda(), w.cuda(), direct_f32=True).cpu()
            want = F.conv2d(xq, w, padding=1)
            err = (y - want).abs().max().item()
            info = {"err": round(err, 5)}
            if err > 1e-3:
                # does it match another tap's result?
                for ky2 in range(3):
                    for kx2 in range(3):
                        w2 = torch.zeros_like(w)
                        w2[:, :, ky2, kx2] = w[:, :, ky, kx]
                        e2 = (y - F.conv2d(xq, w2, padding=1)).abs().max().item()
                        if e2 < 1e-3:
                            info["matches_tap"] = [ky2, kx2]
                bad = (y - want).abs() > 1e-3
                info["bad_frac"] = round(bad.float().mean().item(), 4)
                info["bad_cols"] = sorted(set(bad.nonzero()[:, 3].tolist()))[:12]
                info["bad_rows"] = sorted(set(bad.nonzero()[:, 2].tolist()))[:12]
                info["bad_ch"] = sorted(set(bad.nonzero()[:, 1].tolist()))[:12]
                info["nan"] = bool(torch.isnan(y).any())
            res["taps"][f"{ky}{kx}"] = info
    # full random conv
    w = torch.randn(C, C, 3, 3, generator=g) * 0.05
    b = torch.randn(C, generator=g)
    y = eng.conv3x3(x.cuda(), w.cuda(), b.cuda(), direct_f32=True).cpu()
    want = F.conv2d(xq, w.half().float(), b, padding=1)
    res["full_err"] = round((y - want).abs().max().item(), 5)
    y = eng.conv3x3(x.cuda(), w.cuda(), b.cuda(), direct_f32=False).cpu()
    res["full_err_nhwc"] = round((y - want).abs().max().item(), 5)
    print("DIAG " + json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
        sys.exit(0)
    import __graft_entry__ as g
    g.build()
    for mode in (0, 1, 2):
        env = dict(os.environ, SS4K_DESC_MODE=str(mode), SS4K_SKIP_PROBE="1")
        try:
            p = subprocess.run([sys.executable, __file__, str(mode)], env=env, capture_output=True, text=True, timeout=300)
            out = [l for l in p.stdout.splitlines() if l.startswith("DIAG ")]
            print(f"--- desc mode {mode}: rc={p.returncode}")
            print(out[0] if out else (p.stdout[-1500:] + "\n" + p.stderr[-2500:]))
        except subprocess.TimeoutExpired:
            print(f"--- desc mode {mode}: TIMEOUT")
    # auto-probe
    p = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path.insert(0, %r); import ss4k_b200; e = ss4k_b200.Engine.get(0); print('AUTO desc_mode', e.desc_mode)" % ROOT],
                       capture_output=True, text=True, timeout=300)
    print(p.stdout[-500:], p.stderr[-1500:])
