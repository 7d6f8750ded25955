"""BSVD ring-buffer streaming at 1280x720: ms per pushed frame in steady state, with one graph launch per push
(default) and with the per-layer launches (SS4K_NO_STREAM_GRAPH=1), NV12 frames in, both precision modes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L
from ss4k_b200 import bsvd as nbsvd
from oracle import bsvd as obsvd

H, W, N = 720, 1280, 96
frames = torch.randint(16, 236, (8, H * 3 // 2, W), dtype=torch.uint8, device="cuda")
for split in (False, True):
    sd = obsvd.build_bsvd32(0, weight_scale=1.0 if split else 0.5)
    den = nbsvd.NativeBSVD(sd, device=0, act_mode=L.ACT_F16_SPLIT if split else L.ACT_F16)
    for graph in (True, False):
        if graph:
            os.environ.pop("SS4K_NO_STREAM_GRAPH", None)
        else:
            os.environ["SS4K_NO_STREAM_GRAPH"] = "1"
        s = den.stream(H, W, in_fmt=L.FMT_NV12, noise=0.075)
        for i in range(40):                       # fill the pipeline, capture every phase graph
            s.push(frames[i % 8])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(N):
            s.push(frames[i % 8])
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1000 / N
        torch.cuda.synchronize()
        print(json.dumps({"precision": "split" if split else "f16", "one_graph_per_push": graph,
                          "gpu_ms_per_frame": round(e0.elapsed_time(e1) / N, 3), "host_ms_per_push": round(host_ms, 3),
                          "latency_frames": s.latency}))
        s.close()
