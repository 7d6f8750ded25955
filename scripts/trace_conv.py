"""clock64 phase stamps of the streaming conv kernel (CTA 0, middle, last), cycles since kernel entry:
[0 entry, 1 setup done, 2 weights landed (MMA warp), 3 first slab landed, 4 last MMA issued, 5 first accumulator
complete (epilogue warp 2), 6 last row stored, 7 stores drained, 8 TMEM freed]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ss4k_b200
from bench_conv import bench
eng = ss4k_b200.Engine.get(0)
for cin, cout, n in [(64, 32, 1), (160, 32, 1), (192, 64, 1), (64, 64, 1), (64, 32, 4)]:
    for flags in (0, 6, 7):
        d = bench(eng, cin, cout, 360, 640, n=n, pitch=192 if cout == 32 or cin == 192 else 0, flags=flags, trace=1)
        tr = d.pop("trace")
        print(json.dumps(d))
        for t in tr[:2]:
            print("   phases", t[:9], "setup: barriers", t[9], "tmem", t[10], "bias", t[11], "producer at dep wait", t[12], "after", t[13])
            for i in range(8):
                q = t[16 + 6 * i: 22 + 6 * i]
                print("   row", i, "start", q[0], " +issue8", q[1] - q[0], " +fetch", q[2] - q[1], " +prepare", q[3] - q[2], " +issue4", q[4] - q[3], " +commits", q[5] - q[4])
