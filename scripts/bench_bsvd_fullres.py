"""Pipeline-isolation sweep and clock64 role traces of the BSVD full-resolution convs (720p, 8 frames): which warp role
sets the row period when the MMA work per row is small (flags: 1 no MMA, 2 no TMA loads, 4 no epilogue math/stores);
src = 3: the first layer with the NV12 frames decoded by the kernel's decoder warps."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ss4k_b200
from bench_conv import bench
eng = ss4k_b200.Engine.get(0)
for (cin, cout, src) in [(32, 32, 0), (4, 30, 0), (4, 30, 3), (4, 30, 2)]:
    for flags in (0, 1, 4, 7):
        d = bench(eng, cin, cout, 720, 1280, n=8, act=2, trace=1, flags=flags, src=src)
        tr = d.pop("trace")
        d["src"] = src
        print(json.dumps(d))
        t = tr[1]
        print("   phases", t[:9])
        print("   MMA rows (start, +issue8, +fetch, +prepare, +issue4, +commits)")
        for i in range(8):
            q = t[16 + 6 * i: 22 + 6 * i]
            print("     row", i, q[0], [q[j + 1] - q[j] for j in range(5)])
        print("   producer row starts", t[64:80], "deltas", [t[65 + i] - t[64 + i] for i in range(15)])
        print("   producer row 8: start", t[72], "record done", t[112] - t[72], "a_empty", t[113] - t[112], "issue", t[114] - t[113], "syncwarp", t[115] - t[114])
        if src:
            print("   decoder warp 0 rows (start, +load/convert/store16, next start)", [(t[96 + 2 * i], t[97 + 2 * i] - t[96 + 2 * i]) for i in range(8)])
        print("   epilogue warp 2 rows (wait start, +wait, +ld/init, +math/store)")
        for i in range(8):
            q = t[80 + 4 * i: 84 + 4 * i]
            print("     row", 2 * i, q[0], [q[j + 1] - q[j] for j in range(3)])
