#!/bin/bash
# micro-benchmarks of the 64->64 conv with the library built at several commits (gpurun_tmp_so/libss4k_<commit>.so)
cp sharkshark-4k_b200/csrc/libss4k.so /tmp/keep.so
for so in gpurun_tmp_so/libss4k_*.so; do
  cp $so sharkshark-4k_b200/csrc/libss4k.so
  echo "== $so"
  timeout 120 python - <<'P' 2>&1 | tail -4
import sys
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import ss4k_b200
from bench_conv import bench
eng = ss4k_b200.Engine.get(0)
for (cin, cout, hw, n) in [(64, 64, (360, 640), 4), (64, 64, (1440, 2560), 1), (160, 32, (360, 640), 4), (64, 64, (360, 640), 1)]:
    d = bench(eng, cin, cout, hw[0], hw[1], n=n, pitch=192 if cout == 32 else 0, flags=0)
    print(cin, cout, hw, n, d["ms"], d["tflops"], "a_slots", d.get("a_slots"), flush=True)
P
done
cp /tmp/keep.so sharkshark-4k_b200/csrc/libss4k.so
