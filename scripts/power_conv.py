"""Power attribution of the streaming conv kernel: long runs of one conv with pipeline-isolation flags,
nvidia-smi power / SM clock sampled during the run.  Prints J per launch (power x time)."""
import ctypes, json, os, subprocess, sys, threading, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ss4k_b200
from ss4k_b200 import _lib as L

class Smi:
    def __init__(self):
        self.rows, self.stop = [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)
    def run(self):
        while not self.stop.is_set():
            o = subprocess.run(["nvidia-smi", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip().split(",")
            try: self.rows.append((time.time(), float(o[0]), float(o[1])))
            except Exception: pass
            self.stop.wait(0.05)
    def __enter__(self): self.t.start(); return self
    def __exit__(self, *a): self.stop.set(); self.t.join()

def bench(eng, cin, cout, h, w, n, flags, iters, pitch=192, act=1):
    d = L.ConvDesc(); d.struct_size = ctypes.sizeof(L.ConvDesc)
    d.n, d.h, d.w, d.cin, d.cout, d.mode, d.act = n, h, w, cin, cout, 0, act
    d.alpha, d.beta = 1.0, 0.0
    d.reserved[6] = 1
    ms = ctypes.c_float(); js = ctypes.c_void_p()
    L.check(eng.lib.ss4k_debug_bench_conv(eng.h, ctypes.byref(d), pitch, flags, iters, ctypes.byref(ms), ctypes.byref(js)), eng.h)
    eng.lib.ss4k_free(js)
    return ms.value

if __name__ == "__main__":
    eng = ss4k_b200.Engine.get(0)
    for cin, cout, n in [(160, 32, 4), (64, 32, 4), (192, 64, 4), (160, 32, 1)]:
        base = bench(eng, cin, cout, 360, 640, n, 0, 50)
        for flags in (0, 1, 2, 4, 5, 6, 7):
            est = bench(eng, cin, cout, 360, 640, n, flags, 50)
            iters = int(2500 / max(est, 1e-3))
            with Smi() as s:
                t0 = time.time(); ms = bench(eng, cin, cout, 360, 640, n, flags, iters); t1 = time.time()
            rows = [r for r in s.rows if t0 + 0.8 < r[0] < t1 - 0.1]
            pw = statistics.median(r[1] for r in rows) if rows else -1
            ck = statistics.median(r[2] for r in rows) if rows else -1
            fl = 2.0 * cin * cout * 9 * 360 * 640 * n
            print(json.dumps({"cin": cin, "cout": cout, "n": n, "flags": flags, "us": round(ms * 1000, 2), "tflops": round(fl / ms / 1e9), "power_w": pw, "sm_mhz": ck,
                              "mJ_per_launch": round(pw * ms, 3), "kclk": round(ms * ck, 1), "samples": len(rows)}), flush=True)
            time.sleep(1.0)
