"""Producer-warp statistics of the fused residual-dense-block launches (SS4K_RDB_TRACE=1): polls of the progress
counters, clocks spent waiting for them / for free activation slabs, clock at which every phase starts."""
import ctypes, os, sys
os.environ["SS4K_RDB_TRACE"] = "1"
os.environ["SS4K_RDB_FUSE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ss4k_b200
from ss4k_b200 import _lib as L, realesrgan
from oracle import rrdbnet
torch.manual_seed(0)
net = rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval()
m = realesrgan.NativeRRDBNet(net.state_dict(), scale=2, num_block=23, device=0)
plan = m._plan(1, 720, 1280, L.FMT_U8_NHWC, L.FMT_U8_NHWC)
x = torch.randint(0, 256, (1, 720, 1280, 3), dtype=torch.uint8, device="cuda")
for _ in range(5):
    plan.run(x)
torch.cuda.synchronize()
nf = plan.fused_blocks
cap = nf * 148 * 16
buf = (ctypes.c_longlong * cap)()
n = plan.lib.ss4k_debug_rdb_trace(plan.h, buf, cap)
t = np.frombuffer(buf, dtype=np.int64)[:n].reshape(nf, 148, 16)
t = t[:, :145]
names = ["polls that waited", "clk waiting for counters", "polls", "clk in polls", "clk waiting for free slabs", "producer clk total"]
for g in (1, 30, 60):
    print(f"--- fused launch {g}")
    for i, nm in enumerate(names):
        v = t[g, :, i]
        print(f"  {nm:28s} mean {v.mean():10.0f}  max {v.max():10d}  (cta {int(v.argmax())})")
    ph = t[g, :, 6:12]
    print("  phase start clk (mean over CTAs):", [int(a) for a in ph.mean(axis=0)])
    print("  phase start clk (max  over CTAs):", [int(a) for a in ph.max(axis=0)])
    d = np.diff(np.concatenate([ph, t[g, :, 5:6]], axis=1), axis=1)
    print("  phase length clk (mean):", [int(a) for a in d.mean(axis=0)], " (max):", [int(a) for a in d.max(axis=0)])
    w = t[g, :, 1]
    worst = np.argsort(-w)[:8]
    print("  CTAs waiting longest:", [(int(c), int(w[c]), int(t[g, c, 0])) for c in worst])
