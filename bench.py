#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task prompt / DESIGN.md section 8).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on that fits one GPU):
  RealESRGAN RRDBNet-23 x2 on synthetic 1280x720 frames -> 2560x1440, no denoiser, random-init weights
  (upstream init, seed 0), fp16 operands / fp32 accumulate, `--batch` frames per step.
A step = one pass of the network over one batch of frames.
  value : frames/s, inputs (uint8 NHWC) already resident in HBM, outputs (uint8 NHWC) left in HBM
  e2e   : frames/s through ss4k_run_host_async: pinned host uint8 frames in -> pinned host uint8 frames out,
          H2D and D2H copies inside the timed region
  --impl reference : the oracle's CPU fp32 RRDBNet (the reference's arithmetic lives in pip `basicsr`,
          which is not installable here: "port"), all host threads, a bounded crop per step
N > 1 (torchrun): frames are sharded across ranks (weak scaling, no collective inside the nets); the
uint8 output frames of every rank are gathered to rank 0 (the encoder rank) over NCCL inside the step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAME_H, FRAME_W, SCALE, BLOCKS = 720, 1280, 2, 23
METRIC = "frames/s 720p->1440p RRDBNet x2 (RealESRGAN), no denoiser"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1392.0), d.get("hbm_gbs", 6546.2), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop = gpu_index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w": statistics.median(pw) if pw else None}


def oracle_net():
    from oracle import rrdbnet
    torch.manual_seed(0)
    return rrdbnet.RRDBNet(3, 3, SCALE, 64, BLOCKS, 32).eval()


def cpu_sample(net, crop, reps=1):
    """Times the CPU fp32 path on one crop x crop RGB patch; returns (seconds per patch, frames/s
    extrapolated to a full 1280x720 frame by pixel ratio)."""
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.rand(1, 3, crop, crop, generator=torch.Generator().manual_seed(1234))
    ts = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            net(x)
            ts.append(time.perf_counter() - t0)
    t = min(ts)
    fps = (crop * crop) / (FRAME_H * FRAME_W) / t
    return t, fps


def run_reference(args, rank, world):
    if rank != 0:
        return
    net = oracle_net()
    crop = args.cpu_crop
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_sample(net, crop)
    t0 = time.perf_counter()
    per = [cpu_sample(net, crop)[0] for _ in range(args.steps)]
    total = time.perf_counter() - t0
    fps = (crop * crop) / (FRAME_H * FRAME_W) * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RRDBNet-23 x2 1280x720->2560x1440 (BASELINE.json configs[1])",
                   "sample": f"{crop}x{crop} crop per step, extrapolated by pixel ratio"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"oracle RRDBNet fp32 (restated basicsr arch; pip basicsr not installable), "
                                   f"{crop}x{crop} crop x {args.steps} steps, extrapolated to 1280x720 by pixel ratio"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_native(args, rank, world, local_rank):
    import ss4k_b200
    from ss4k_b200 import _lib as L
    from ss4k_b200 import realesrgan

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    net = oracle_net()
    act_mode = L.ACT_BF16 if args.dtype == "bf16" else L.ACT_F16
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=SCALE, num_block=BLOCKS, device=local_rank, act_mode=act_mode)
    plan = model._plan(B, FRAME_H, FRAME_W, L.FMT_U8_NHWC, L.FMT_U8_NHWC)
    eng = model.engine
    g = torch.Generator().manual_seed(1234 + rank)
    n_in = 3  # rotate inputs
    frames_host = [torch.randint(0, 256, (B, FRAME_H, FRAME_W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_in)]
    frames_dev = [f.to(dev) for f in frames_host]
    out_dev = plan.new_output()
    out_host = torch.empty(plan.out_shape(), dtype=torch.uint8).pin_memory()
    gather_list = None
    if dist is not None and rank == 0:
        gather_list = [torch.empty_like(out_dev) for _ in range(world)]

    def step(i):
        plan.run(frames_dev[i % n_in], out_dev)
        if dist is not None:
            dist.gather(out_dev, gather_list, dst=0)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        for i in range(args.steps):
            step(i)
        ev1.record()
        sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    fps = world * B * args.steps / (ms / 1000)

    # ---- e2e: host frames in, host frames out, copies inside the timed region
    # (ss4k_run_host_async: every step's H2D copy, kernels and D2H copy are queued inside the timed region; the copies
    #  of neighbouring steps overlap the kernels, as in the reference's producer / consumer queues)
    out_host2 = torch.empty_like(out_host).pin_memory()
    outs = (out_host, out_host2)
    for i in range(min(2, args.warmup)):
        plan.run_host_async(frames_host[i % n_in], outs[i & 1])
    plan.host_sync()
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        plan.run_host_async(frames_host[i % n_in], outs[i & 1])
    plan.host_sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_fps = world * B * args.steps / e2e_s

    # ---- dominant kernel, measured live: CUDA events between the plan's steps (no graph), same inputs
    prof = []
    for i in range(min(3, args.steps)):
        prof.append(plan.profile(frames_dev[i % n_in], out_dev))
    torch.cuda.synchronize()

    # ---- HBM-bound layout / colour kernels: achieved GB/s against the measured copy bandwidth
    hbm_kernels = []
    if rank == 0:
        import ctypes
        def timed(fn, iters=20):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        hr = out_dev if out_dev.dim() == 4 else out_dev.reshape(B, 2 * FRAME_H, 2 * FRAME_W, 3)
        t = timed(lambda: eng.rgb_to_nv12(hr))
        nbytes = hr.numel() * 1.5
        hbm_kernels.append({"kernel": "rgb_to_nv12_kernel", "bytes_per_launch": nbytes, "us": 1000 * t, "GB/s": nbytes / t / 1e6,
                            "what": "uint8 RGB 2560x1440 -> NV12: 3 B/px read + 1.5 B/px written"})
        big = torch.randint(0, 256, (8, 2 * FRAME_H, 2 * FRAME_W, 3), dtype=torch.uint8, device=dev)
        t8 = timed(lambda: eng.rgb_to_nv12(big))
        hbm_kernels.append({"kernel": "rgb_to_nv12_kernel", "bytes_per_launch": big.numel() * 1.5, "us": 1000 * t8,
                            "GB/s": big.numel() * 1.5 / t8 / 1e6, "what": "the same on 8 frames per launch (133 MB: above launch latency and L2)"})
        del big
        layout = [(ms_, kd) for (ms_, fl, kd) in prof[0] if kd == 0]
        if layout:
            # prep_kernel: uint8 NHWC 1280x720 -> fp16 NHWC, pixel-unshuffle(2), 16-channel pitch: 3 B/px in, 32 B per trunk px out
            nb = B * FRAME_H * FRAME_W * 3 + B * (FRAME_H // 2) * (FRAME_W // 2) * 16 * 2
            tms = sum(m for m, _ in layout) / len(layout)
            hbm_kernels.append({"kernel": "prep_kernel<u8 NHWC>", "bytes_per_launch": nb, "us": 1000 * tms, "GB/s": nb / tms / 1e6,
                                "what": "uint8 frame -> /255 -> pixel_unshuffle(2) -> fp16 NHWC (launch-latency bound at this size)"})

    if rank == 0:
        peak_tf, peak_hbm, peak_src = measured_peaks()
        k_ms = sum(ms for run in prof for (ms, fl, kd) in run if kd == 1) / len(prof)
        k_fl = sum(fl for (ms, fl, kd) in prof[0] if kd == 1)
        k_n = sum(1 for (ms, fl, kd) in prof[0] if kd == 1)
        all_ms = sum(ms for run in prof for (ms, fl, kd) in run) / len(prof)
        other = {"tile_conv_ms": sum(ms for (ms, fl, kd) in prof[0] if kd == 2), "layout_ms": sum(ms for (ms, fl, kd) in prof[0] if kd == 0)}
        share = k_ms / all_ms
        # the kernel's average launch duration inside the TIMED region (graph replay, dependent launches overlapped):
        # timed-region device time x the kernel's share of a step / its launches in the region
        k_us_timed = 1000.0 * (ms / args.steps) * share / max(1, k_n)
        achieved = (k_fl / max(1, k_n)) / (k_us_timed * 1e-6) / 1e12
        achieved_events = k_fl / (k_ms / 1000) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_v7_dram_traffic_b1.json")
        if B == 1 and os.path.isfile(tp):
            with open(tp) as f:
                tk = json.load(f)["kernels"]
            tot_b = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in tk.items() if "conv3x3_stream" in k)
            tot_n = sum(v["launches"] for k, v in tk.items() if "conv3x3_stream" in k)
            traffic = tot_b / max(1, tot_n)
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if act_mode == L.ACT_F16 else "bf16", "data": "synthetic",
            "config": {"workload": "RRDBNet-23 x2 1280x720->2560x1440 (BASELINE.json configs[1])",
                       "frames_per_step_per_gpu": B, "in": "uint8 NHWC", "out": "uint8 NHWC",
                       "weights": "random init (upstream basicsr init, seed 0)",
                       "l2": "no flush: one frame reads and writes 118 MB slabs 69 times (13.5 GB through HBM per frame, >> 126 MB L2); 3 input buffers rotated",
                       "desc_mode": eng.desc_mode, "parallelism": f"frame-sharded x{world}",
                       "launch": "%d of %d kernels per step replay from one CUDA graph; programmatic dependent launch %s"
                                 % (plan.graph_steps, plan.launches, "off" if os.environ.get("SS4K_NO_PDL") else "on")},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": plan.in_bytes, "d2h_bytes_per_step": plan.out_bytes},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "hbm_kernels": {"peak_GB/s": peak_hbm, "kernels": hbm_kernels},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "kernel": "conv3x3_stream_kernel", "peak_source": peak_src,
                         "launches_per_step": k_n, "avg_launch_us": k_us_timed,
                         "flops_per_launch_avg": k_fl / max(1, k_n), "kernel_share_of_step": share,
                         "achieved_unoverlapped": achieved_events, "avg_launch_us_unoverlapped": 1000 * k_ms / max(1, k_n),
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the kernel's launches of one frame "
                                         "(profiles/r01_v7_dram_traffic_b1.json; one ncu pass, warm caches, batch 1)" if traffic else None,
                         "how": "avg launch duration = CUDA-event time of the timed region (graph replay) x the kernel's share of a step / "
                                "its launches; share from CUDA events between every step of an un-graphed run on the launching stream "
                                "(mean of %d runs; those serialised per-launch times give 'achieved_unoverlapped'); "
                                "algorithmic FLOPs = 2*Cin*Cout*9*Hout*Wout" % len(prof),
                         "bound_note": "tensor pipe, fed from shared memory: an M=128,N=96,K=16 MMA reads 7 KB of operands = 56 clk at "
                                       "128 B/clk/SM vs 48 clk of math (scripts/mma_issue_probe.cu), see DESIGN.md section 4",
                         "whole_step_tflops": plan.flops * args.steps / (ms / 1000) / 1e12, **other},
        }
        if world == 1 and not args.no_cpu:
            t, cfps = cpu_sample(net, args.cpu_crop)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"oracle RRDBNet fp32, one {args.cpu_crop}x{args.cpu_crop} crop ({t:.1f} s), extrapolated to 1280x720 by pixel ratio"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--dtype", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--cpu-crop", type=int, default=384)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        args.steps = min(args.steps, 12)   # each CPU step is ~1 s: keep the reference arm within minutes
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    so = os.path.join(ROOT, "sharkshark-4k_b200", "csrc", "libss4k.so")
    if rank == 0 and (world == 1 or not os.path.isfile(so)):
        import __graft_entry__ as g
        g.build()
    t0 = time.time()
    while not os.path.isfile(so) and time.time() - t0 < 600:
        time.sleep(1.0)
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
