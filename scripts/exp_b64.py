import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
import ss4k_b200
from bench_conv import bench
eng = ss4k_b200.Engine.get(0)
for (cin, cout, hw, n) in [(64, 64, (360, 640), 4), (64, 64, (1440, 2560), 1), (160, 32, (360, 640), 4), (64, 64, (360, 640), 1), (64, 32, (360, 640), 1)]:
    d = bench(eng, cin, cout, hw[0], hw[1], n=n, pitch=192 if cout == 32 else 0, flags=0)
    print(cin, cout, hw, n, d["ms"], d["tflops"], "a_slots", d.get("a_slots"), flush=True)
