"""clock64 stamps of epilogue warp 2 (dbg flag 64) for its first 8 rows: wait for the accumulator, tcgen05.ld, slot
re-initialisation (tcgen05.st), activation / pack / staging stores, TMA store issue."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ss4k_b200
from bench_conv import bench
eng = ss4k_b200.Engine.get(0)
for cin, cout, hw, n in [(64, 64, (1440, 2560), 1), (64, 64, (360, 640), 4), (160, 32, (360, 640), 4)]:
    d = bench(eng, cin, cout, hw[0], hw[1], n=n, pitch=192 if cout == 32 else 0, flags=64, trace=1)
    tr = d.pop("trace")
    print(json.dumps(d))
    t = tr[0]
    print("   phases", t[:9])
    for i in range(8):
        q = t[16 + 6 * i: 22 + 6 * i]
        print("   epilogue row", 2 * i, "start", q[0], " +acc wait", q[1] - q[0], " +tmem ld", q[2] - q[1], " +re-init", q[3] - q[2], " +math+sts", q[4] - q[3], " +tma issue", q[5] - q[4])
