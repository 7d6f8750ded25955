#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task prompt / DESIGN.md section 8).

Default workload = BASELINE.json configs[2] (the configuration the metric "frames/s 720p->1440p denoise+SR" is quoted
on; it fits one GPU):
  synthetic NV12 1280x720 stream -> BSVD-32 temporal denoiser over a chunk of `--clip` frames (reference constructor
  init -> fp16 hi/lo split precision, the parity configuration) -> sharpen / clamp / 0.8-0.2 blend with the decoded
  frame (fsrcnn_upscaler.py:278-281) -> RRDBNet-23 x2 on every owned frame -> uint8 RGB 2560x1440.
  A step = one chunk of `--clip` owned frames per GPU.
  N > 1 (torchrun): the global clip of N*clip frames is sharded into contiguous chunks with the denoiser's 16-frame
  temporal halo (sharding.bsvd_chunks): halo frames are decoded and denoised redundantly, only owned frames are
  upscaled; the finished uint8 frames are gathered to rank 0 (the encoder rank) over NCCL inside every step.
  value : owned frames/s of the whole job, NV12 chunks already resident in HBM, uint8 frames left in HBM (rank 0 after
          the gather)
  e2e   : the same through the public host path: pinned host NV12 chunk in -> pinned host uint8 frames out (on rank 0,
          after the gather), H2D and D2H copies inside the timed region
`--workload cfg2` keeps the SR-only line (RRDBNet-23 x2, uint8 frames, no denoiser: BASELINE.json configs[1]).
`--impl reference`: the reference's arithmetic on the host cores (oracle port: basicsr is not installable here), one
full 1280x720 frame per step.
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
DENOISE_RATE = 0.75                      # CLI default of the reference, src/main/upscaler.py:25
NOISE = 0.1 * DENOISE_RATE               # fsrcnn_upscaler.py:262
METRIC_CFG3 = "frames/s 720p->1440p denoise+SR (BSVD-32 + RealESRGAN RRDBNet x2)"
METRIC_CFG2 = "frames/s 720p->1440p RRDBNet x2 (RealESRGAN), no denoiser"
WORKLOAD_CFG3 = "BSVD denoise + RRDBNet-23 x2, 720p NV12 stream -> 1440p uint8 RGB (BASELINE.json configs[2])"
WORKLOAD_CFG2 = "RRDBNet-23 x2 1280x720->2560x1440 (BASELINE.json configs[1])"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops_sustained", 1392.0), d.get("bf16_tflops", 1650.0), d.get("hbm_gbs", 6546.2),
                "measured (MEASURED_PEAKS.json, sustained bf16; the burst figure is peak_burst)")
    return 1400.0, 1650.0, 6650.0, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------- weights / data
def oracle_rrdb():
    """Seeded random-init RRDBNet (upstream init); the module also serves the CPU legs."""
    from oracle import rrdbnet
    torch.manual_seed(0)
    return rrdbnet.RRDBNet(3, 3, SCALE, 64, BLOCKS, 32).eval()


def bsvd_state(weight_scale=1.0):
    """BSVD-32 with the reference constructor's init (model.py:393-400,501-508), seed 0."""
    from oracle import bsvd
    return bsvd.build_bsvd32(0, weight_scale=weight_scale)


def synth_nv12(t0, t1, device, variant=0):
    """Synthetic NV12 720p frames [t1-t0, H*W*3/2]: smooth moving pattern + N(0, 10 LSB) noise on luma; the frame with
    global index t is the same on every rank."""
    H, W = FRAME_H, FRAME_W
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :]
    out = torch.empty(t1 - t0, H * W * 3 // 2, dtype=torch.uint8, device=device)
    for i, t in enumerate(range(t0, t1)):
        g = torch.Generator(device=device).manual_seed(1234 + 7919 * variant + t)
        y = 126 + 70 * torch.sin((xx + 4 * t) * 0.013) * torch.cos((yy + 2 * t) * 0.017) + 30 * torch.sin((xx + yy) * 0.05)
        y = y + 10 * torch.randn(H, W, device=device, generator=g)
        out[i, :H * W] = y.clamp(16, 235).round().to(torch.uint8).reshape(-1)
        u = 128 + 50 * torch.sin((xx[:, ::2] + 3 * t) * 0.011) + 0 * yy[::2]
        v = 128 + 50 * torch.cos((yy[::2] + 5 * t) * 0.009) + 0 * xx[:, ::2]
        uv = torch.stack([u, v], dim=-1).clamp(16, 240).round().to(torch.uint8)
        out[i, H * W:] = uv.reshape(-1)
    return out


# ---------------------------------------------------------------------------------------------- CPU legs
def cpu_frame_cfg3(net, bsd, crop_h=FRAME_H):
    """One frame of cfg3 on the host cores in fp32: NV12 decode -> BSVD (one-frame clip) -> RRDBNet x2.  Returns seconds.
    crop_h < 720 times a band of the frame (bounded sample), full width."""
    from oracle import bsvd, colour, glue
    torch.set_num_threads(os.cpu_count() or 1)
    nv = synth_nv12(0, 1, torch.device("cpu")).numpy()
    t0 = time.perf_counter()
    with torch.no_grad():
        rgb = torch.from_numpy(colour.nv12_to_rgb(nv, FRAME_H, FRAME_W))[:, :, :crop_h]
        x = torch.cat([rgb, torch.full((1, 1, crop_h, FRAME_W), NOISE)], dim=1)[None]
        den = bsvd.bsvd_forward(bsd, x)[0]
        den = torch.clamp(glue.depthwise_reflect(den.reshape(3, 1, crop_h, FRAME_W), glue.sharpen_weight(0.00002)).reshape(1, 3, crop_h, FRAME_W), 0, 1)
        hr = net(den * 0.8 + 0.2 * rgb)                 # fsrcnn_upscaler.py:278-281
        (hr.clamp(0, 1) * 255).to(torch.uint8)
    return time.perf_counter() - t0


def cpu_frame_cfg2(net, crop_h=FRAME_H):
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.rand(1, 3, crop_h, FRAME_W, generator=torch.Generator().manual_seed(1234))
    t0 = time.perf_counter()
    with torch.no_grad():
        net(x)
    return time.perf_counter() - t0


def cpu_baseline(workload, budget_s=45.0):
    """1 warm-up + up to 3 timed full frames, bounded by `budget_s` of CPU work."""
    net = oracle_rrdb()
    bsd = bsvd_state() if workload == "cfg3" else None
    fn = (lambda: cpu_frame_cfg3(net, bsd)) if workload == "cfg3" else (lambda: cpu_frame_cfg2(net))
    t_start = time.perf_counter()
    fn()
    ts = []
    while len(ts) < 3 and (not ts or time.perf_counter() - t_start + ts[-1] < budget_s):
        ts.append(fn())
    t = statistics.median(ts)
    what = ("oracle BSVD-32 (reference constructor init, one-frame clip) + oracle RRDBNet-23 x2" if workload == "cfg3"
            else "oracle RRDBNet-23 x2")
    return {"value": 1.0 / t, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{what}, fp32 torch CPU, full 1280x720 frame: 1 warm-up + median of {len(ts)} ({t:.2f} s per frame); "
                      "RRDBNet arithmetic is a restatement of pip basicsr (not installable offline)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    net = oracle_rrdb()
    bsd = bsvd_state() if args.workload == "cfg3" else None
    cores = os.cpu_count() or 1
    crop_h = FRAME_H
    one = (lambda: cpu_frame_cfg3(net, bsd, crop_h)) if args.workload == "cfg3" else (lambda: cpu_frame_cfg2(net, crop_h))
    t_probe = one()                                   # first call (also a warm-up)
    # keep the whole run within a few minutes: a band of the frame when a full frame is too slow for steps + warmup
    total = args.steps + args.warmup
    while t_probe * total * (crop_h / FRAME_H) > 420.0 and crop_h > 180:
        crop_h //= 2
    scale = crop_h / FRAME_H
    for _ in range(max(0, args.warmup - 1)):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    total_s = time.perf_counter() - t0
    fps = scale * args.steps / total_s
    sample = ("full 1280x720 frame per step" if crop_h == FRAME_H else
              f"1280x{crop_h} band per step, extrapolated to 1280x720 by pixel ratio")
    what = ("oracle BSVD-32 (one-frame clip) + oracle RRDBNet-23 x2" if args.workload == "cfg3" else "oracle RRDBNet-23 x2")
    line = {
        "impl": "reference", "metric": METRIC_CFG3 if args.workload == "cfg3" else METRIC_CFG2, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_CFG3 if args.workload == "cfg3" else WORKLOAD_CFG2, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{what}, fp32 torch CPU (restated basicsr arch; pip basicsr not installable), {sample} x {args.steps} steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- native helpers
def timed_kernel(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def dram_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the streaming conv kernel from the newest committed
    single-pass ncu capture (profiles/r*_dram_traffic_b1.json), or None."""
    pd = os.path.join(ROOT, "profiles")
    cands = sorted(f for f in os.listdir(pd) if f.endswith("_dram_traffic_b1.json")) if os.path.isdir(pd) else []
    if not cands:
        return None, None
    with open(os.path.join(pd, cands[-1])) as f:
        tk = json.load(f)["kernels"]
    keys = [k for k in tk if "conv3x3_stream" in k or "rdb_fused" in k]
    tot_b = sum(tk[k]["dram_read_bytes"] + tk[k]["dram_write_bytes"] for k in keys)
    tot_n = sum(tk[k]["launches"] for k in keys)
    return (tot_b / tot_n if tot_n else None), cands[-1]


def roofline_block(prof_runs, weights, step_ms, whole_flops_per_step):
    """prof_runs: list of (profile rows [(ms, flops, kind)...], multiplicity per step).  The dominant kernel is the
    row-streaming tcgen05 conv kernel (kind 1 / 3)."""
    peak_tf, peak_burst, peak_hbm, peak_src = measured_peaks()
    conv_kinds = (1, 3)
    k_ms = sum(mult * sum(ms for (ms, fl, kd) in rows if kd in conv_kinds) for rows, mult in prof_runs)
    k_fl = sum(mult * sum(fl for (ms, fl, kd) in rows if kd in conv_kinds) for rows, mult in prof_runs)
    k_n = sum(mult * sum(1 for (ms, fl, kd) in rows if kd in conv_kinds) for rows, mult in prof_runs)
    all_ms = sum(mult * sum(ms for (ms, fl, kd) in rows) for rows, mult in prof_runs)
    tile_ms = sum(mult * sum(ms for (ms, fl, kd) in rows if kd == 2) for rows, mult in prof_runs)
    layout_ms = sum(mult * sum(ms for (ms, fl, kd) in rows if kd == 0) for rows, mult in prof_runs)
    share = k_ms / all_ms if all_ms else 0.0
    k_us_timed = 1000.0 * step_ms * share / max(1, k_n)
    achieved = (k_fl / max(1, k_n)) / (k_us_timed * 1e-6) / 1e12
    traffic, tsrc = dram_traffic_per_launch()
    return {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": traffic, "kernel": "conv3x3_stream_kernel (row-streaming tcgen05 implicit GEMM)",
            "peak_source": peak_src, "peak_burst": peak_burst, "frac_of_burst": achieved / peak_burst,
            "frac_of_nominal_2250": achieved / 2250.0,
            "launches_per_step": k_n, "avg_launch_us": k_us_timed, "flops_per_launch_avg": k_fl / max(1, k_n),
            "kernel_share_of_step": share, "achieved_unoverlapped": k_fl / (k_ms / 1000) / 1e12 if k_ms else None,
            "traffic_note": (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the kernel's launches of one "
                             f"RRDBNet frame (profiles/{tsrc}; one ncu pass, warm caches)") if traffic else None,
            "how": "avg launch duration = CUDA-event time of the timed region x the kernel's share of a step / its launches per "
                   "step; share from CUDA events between every step of un-graphed plan runs on the launching stream; algorithmic "
                   "FLOPs = 2*Cin*Cout*9*Hout*Wout, true channel counts",
            "bound_note": "tensor pipe fed from shared memory: an M=128,N=96,K=16 MMA reads 7 KB of operands = 56 clk at 128 B/clk/SM "
                          "vs 48 clk of math (scripts/mma_issue_probe.cu), DESIGN.md section 4",
            "whole_step_tflops": whole_flops_per_step / (step_ms / 1000) / 1e12,
            "tile_conv_ms_per_step": tile_ms, "layout_ms_per_step": layout_ms}


# ---------------------------------------------------------------------------------------------- cfg3
def run_native_cfg3(args, rank, world, local_rank):
    import ss4k_b200
    from ss4k_b200 import _lib as L
    from ss4k_b200 import bsvd as nb, realesrgan, sharding
    from ss4k_b200.pipeline import DenoiseUpscalePipeline

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    F = args.clip
    T_all = world * F
    ch = sharding.bsvd_chunks(T_all, world)[rank]
    T = ch.load_hi - ch.load_lo
    rr = oracle_rrdb()
    act_sr = L.ACT_BF16 if args.dtype == "bf16" else L.ACT_F16
    sr = realesrgan.NativeRRDBNet(rr.state_dict(), scale=SCALE, num_block=BLOCKS, device=local_rank, act_mode=act_sr)
    bsvd_mode = {"split": L.ACT_F16_SPLIT, "f16": L.ACT_F16, "auto": "auto"}[args.bsvd]
    den = nb.NativeBSVD(bsvd_state(1.0 if args.bsvd != "f16" else 0.5), device=local_rank, act_mode=bsvd_mode,
                        out_dtype=torch.float16)
    pipe = DenoiseUpscalePipeline(den, sr, FRAME_H, FRAME_W, NOISE, nv12=True, out_fmt=L.FMT_U8_NHWC)
    eng = sr.engine
    n_in = 2
    chunks_dev = [synth_nv12(ch.load_lo, ch.load_hi, dev, variant=v) for v in range(n_in)]
    chunks_host = [c.cpu().pin_memory() for c in chunks_dev]
    out_dev = pipe.new_output(F)
    gather_list = [torch.empty_like(out_dev) for _ in range(world)] if (dist is not None and rank == 0) else None

    def step(i):
        pipe.run(chunks_dev[i % n_in], ch.owned, out_dev)
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
    fps = T_all * args.steps / (ms / 1000)

    # ---- e2e: pinned host NV12 chunk in -> pinned host uint8 frames out (rank 0, after the gather)
    if dist is None:
        outs_host = [torch.empty((F,) + pipe.out_frame_shape(), dtype=torch.uint8).pin_memory() for _ in range(2)]

        def e2e_step(i):
            pipe.run_host(chunks_host[i % n_in], ch.owned, outs_host[i & 1])

        def e2e_sync():
            pipe.host_sync()
        d2h_bytes = outs_host[0].numel()
    else:
        stage_in = [torch.empty_like(chunks_dev[0]) for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        host_all = [torch.empty((world, F) + pipe.out_frame_shape(), dtype=torch.uint8).pin_memory() for _ in range(2)] if rank == 0 else None
        # two sets of gather buffers on the encoder rank: the D2H copy of step i overlaps the kernels AND the gather of step i+1
        gather_lists = [gather_list, [torch.empty_like(out_dev) for _ in range(world)]] if rank == 0 else [None, None]
        copied = [None, None]

        def e2e_step(i):
            s = i & 1
            cur = torch.cuda.current_stream(dev)
            stage_in[s].copy_(chunks_host[i % n_in], non_blocking=True)
            pipe.run(stage_in[s], ch.owned, out_dev)
            if rank == 0 and copied[s] is not None:
                cur.wait_event(copied[s])     # slot s's frames of step i-2 have left the device
            dist.gather(out_dev, gather_lists[s], dst=0)
            if rank == 0:
                ev = torch.cuda.Event()
                ev.record(cur)
                copy_stream.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    for r in range(world):
                        host_all[s][r].copy_(gather_lists[s][r], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                copied[s] = done

        def e2e_sync():
            torch.cuda.synchronize()
        d2h_bytes = world * F * out_dev[0].numel() if rank == 0 else 0
    for i in range(min(2, args.warmup)):
        e2e_step(i)
    e2e_sync()
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    e2e_sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_fps = T_all * args.steps / e2e_s

    # ---- per-kernel-class shares: un-graphed plan runs with CUDA events between the steps
    den_plan = pipe.den_plan(T, ch.lo - ch.load_lo, ch.hi - ch.load_lo)
    den_out = den_plan.new_output()
    prof_den = den_plan.profile(chunks_dev[0], den_out)
    lr0 = torch.rand(1, 3, FRAME_H, FRAME_W, device=dev)
    prof_sr = pipe.sr_plan.profile(lr0, out_dev[0:1])
    prof_sr = pipe.sr_plan.profile(lr0, out_dev[0:1])
    torch.cuda.synchronize()
    glue_ms = None
    if pipe._act is not None:
        # the product path never runs the upscaler plan's layout step: the glue between the nets (sharpen + clamp + blend
        # with the decoded NV12 frame) writes conv_first's activation tensor itself -- time that kernel in its place
        import ctypes
        ptr, pitch, us, bf = pipe._act
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        glue_ms = timed_kernel(lambda: eng.lib.ss4k_glue_sharpen_blend_act(
            ctypes.c_void_p(den_out.data_ptr()), 1, 1, 3, FRAME_H, FRAME_W, 0.00002, 0.8, ctypes.c_void_p(chunks_dev[0].data_ptr()), 3,
            ctypes.c_void_p(ptr), us, pitch, bf, st))
        prof_sr = [(glue_ms, 0.0, 0) if i == 0 and kd == 0 else (m, f, kd) for i, (m, f, kd) in enumerate(prof_sr)]
    den_ms = sum(m for m, _, _ in prof_den)
    sr_ms = sum(m for m, _, _ in prof_sr)

    if rank == 0:
        step_ms = ms / args.steps
        flops_step = den_plan.flops + F * pipe.sr_plan.flops
        roof = roofline_block([(prof_den, 1), (prof_sr, F)], None, step_ms, flops_step)
        peak_hbm = measured_peaks()[2]
        hbm_kernels = []
        lay = [m for m, _, kd in prof_den if kd == 0]
        if lay:
            nb_ = T * (FRAME_H * FRAME_W * 3 // 2) + T * FRAME_H * FRAME_W * 16 * 2 * (2 if den.act_mode == L.ACT_F16_SPLIT else 1)
            hbm_kernels.append({"kernel": "prep_kernel<NV12>", "bytes_per_launch": nb_, "us": 1000 * lay[0], "GB/s": nb_ / lay[0] / 1e6,
                                "what": f"NV12 -> RGB (BT.709) + noise map -> 16-channel fp16 NHWC ({'hi + lo twins, ' if den.act_mode == L.ACT_F16_SPLIT else ''}{T} frames per launch): 1.5 B/px read, 32 B/px written per twin"})
        if glue_ms is not None:
            nb_ = FRAME_H * FRAME_W * (3 * 2 + 1.5 + 8)
            hbm_kernels.append({"kernel": "sharpen_blend_act_us2_kernel", "bytes_per_launch": nb_, "us": 1000 * glue_ms, "GB/s": nb_ / glue_ms / 1e6,
                                "what": "glue between the nets, one frame per launch: 3x3 reflect sharpen + clamp of the denoised frame (fp16 NCHW, 6 B/px), "
                                        "0.8/0.2 blend with the NV12 frame decoded in place (1.5 B/px) -> RRDBNet conv_first's pixel-unshuffled fp16 NHWC input (8 B/px)"})
        t_nv = timed_kernel(lambda: eng.rgb_to_nv12(out_dev[:8]))
        nbytes = out_dev[:8].numel() * 1.5
        hbm_kernels.append({"kernel": "rgb_to_nv12_kernel", "bytes_per_launch": nbytes, "us": 1000 * t_nv, "GB/s": nbytes / t_nv / 1e6,
                            "what": "uint8 RGB 2560x1440 -> NV12 (encoder side), 8 frames per launch: 3 B/px read + 1.5 B/px written"})
        line = {
            "metric": METRIC_CFG3, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if act_sr == L.ACT_F16 else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD_CFG3,
                       "frames_per_step_per_gpu": F, "frames_denoised_per_step_this_rank": T,
                       "in": "NV12 (BT.709 limited range)", "out": "uint8 RGB NHWC",
                       "bsvd_precision": {L.ACT_F16_SPLIT: "fp16 hi/lo split (3 MMAs per product): the parity configuration for the reference constructor init",
                                          L.ACT_F16: "fp16 single MMA (trained-like weights: constructor init x 0.5)"}.get(den.act_mode, str(den.act_mode)),
                       "weights": "random init: BSVD reference constructor (kaiming_normal_), RRDBNet upstream basicsr init, seed 0",
                       "noise_map": NOISE,
                       "l2": "no flush: every step streams > 100 GB through HBM (BSVD clip tensors of %d frames, 13.5 GB per upscaled frame), >> 126 MB L2; %d input chunks rotated" % (T, n_in),
                       "parallelism": f"contiguous frame chunks x{world}, 16-frame BSVD halo per side (sharding.bsvd_chunks; every BSVD layer runs only on the halo frames the owned outputs depend on), NCCL gather of uint8 frames to rank 0 inside the step",
                       "ms_per_frame": {"bsvd_per_owned_frame": den_ms / F, "rrdb_per_upscaled_frame": sr_ms,
                                        "note": "un-graphed profile runs (serialised launches)"},
                       "colour": "NV12 -> RGB (BT.709), /255 and the noise map are decoded by the first BSVD conv's loader warps (no layout kernel); "
                                 "the sharpen / blend glue between the nets writes RRDBNet's first activation tensor; the last conv stores uint8 RGB",
                       "launch": "BSVD clip: %d kernels per chunk, RRDBNet: %d per frame (+ 1 glue kernel, - its layout kernel), CUDA graphs; programmatic dependent launch %s"
                                 % (den_plan.launches, pipe.sr_plan.launches, "off" if os.environ.get("SS4K_NO_PDL") else "on")},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": chunks_host[0].numel(), "d2h_bytes_per_step": d2h_bytes,
                    "note": "per rank: pinned NV12 chunk H2D -> BSVD -> RRDBNet -> (N>1: NCCL gather ->) D2H of the uint8 frames on rank 0"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "timed_region_s": ms / 1000,
            "hbm_kernels": {"peak_GB/s": peak_hbm, "kernels": hbm_kernels},
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline("cfg3")
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- cfg2 (SR only)
def run_native_cfg2(args, rank, world, local_rank):
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
    net = oracle_rrdb()
    act_mode = L.ACT_BF16 if args.dtype == "bf16" else L.ACT_F16
    model = realesrgan.NativeRRDBNet(net.state_dict(), scale=SCALE, num_block=BLOCKS, device=local_rank, act_mode=act_mode)
    plan = model._plan(B, FRAME_H, FRAME_W, L.FMT_U8_NHWC, L.FMT_U8_NHWC)
    eng = model.engine
    g = torch.Generator().manual_seed(1234 + rank)
    n_in = 3
    frames_host = [torch.randint(0, 256, (B, FRAME_H, FRAME_W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_in)]
    frames_dev = [f.to(dev) for f in frames_host]
    out_dev = plan.new_output()
    gather_list = [torch.empty_like(out_dev) for _ in range(world)] if (dist is not None and rank == 0) else None

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

    # ---- e2e: pinned host frames in, pinned host frames out (ss4k_run_host_async); N > 1: + gather + rank-0 D2H
    if dist is None:
        outs = [torch.empty(plan.out_shape(), dtype=torch.uint8).pin_memory() for _ in range(2)]

        def e2e_step(i):
            plan.run_host_async(frames_host[i % n_in], outs[i & 1])

        def e2e_sync():
            plan.host_sync()
        d2h = plan.out_bytes
    else:
        stage = [torch.empty_like(frames_dev[0]) for _ in range(2)]
        host_all = torch.empty((world,) + tuple(out_dev.shape), dtype=torch.uint8).pin_memory() if rank == 0 else None

        def e2e_step(i):
            stage[i & 1].copy_(frames_host[i % n_in], non_blocking=True)
            plan.run(stage[i & 1], out_dev)
            dist.gather(out_dev, gather_list, dst=0)
            if rank == 0:
                for r in range(world):
                    host_all[r].copy_(gather_list[r], non_blocking=True)

        def e2e_sync():
            torch.cuda.synchronize()
        d2h = world * plan.out_bytes if rank == 0 else 0
    for i in range(min(2, args.warmup)):
        e2e_step(i)
    e2e_sync()
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    e2e_sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_fps = world * B * args.steps / e2e_s

    prof = [plan.profile(frames_dev[i % n_in], out_dev) for i in range(2)]
    torch.cuda.synchronize()
    if rank == 0:
        roof = roofline_block([(prof[-1], 1)], None, ms / args.steps, plan.flops)
        peak_hbm = measured_peaks()[2]
        hbm_kernels = []
        hr = out_dev.reshape(B, 2 * FRAME_H, 2 * FRAME_W, 3)
        t_nv = timed_kernel(lambda: eng.rgb_to_nv12(hr))
        hbm_kernels.append({"kernel": "rgb_to_nv12_kernel", "bytes_per_launch": hr.numel() * 1.5, "us": 1000 * t_nv,
                            "GB/s": hr.numel() * 1.5 / t_nv / 1e6, "what": "uint8 RGB 2560x1440 -> NV12: 3 B/px read + 1.5 B/px written"})
        lay = [m for (m, fl, kd) in prof[-1] if kd == 0]
        if lay:
            nb_ = B * FRAME_H * FRAME_W * 3 + B * (FRAME_H // 2) * (FRAME_W // 2) * 16 * 2
            hbm_kernels.append({"kernel": "prep_kernel<u8 NHWC>", "bytes_per_launch": nb_, "us": 1000 * lay[0], "GB/s": nb_ / lay[0] / 1e6,
                                "what": "uint8 frame -> /255 -> pixel_unshuffle(2) -> fp16 NHWC (launch-latency bound at this size)"})
        line = {
            "metric": METRIC_CFG2, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if act_mode == L.ACT_F16 else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD_CFG2, "frames_per_step_per_gpu": B, "in": "uint8 NHWC", "out": "uint8 NHWC",
                       "weights": "random init (upstream basicsr init, seed 0)",
                       "l2": "no flush: one frame streams > 10 GB through HBM (>> 126 MB L2); 3 input buffers rotated",
                       "parallelism": f"frame-sharded x{world}, NCCL gather of uint8 frames to rank 0 inside the step",
                       "launch": "%d of %d kernels per step replay from one CUDA graph; programmatic dependent launch %s"
                                 % (plan.graph_steps, plan.launches, "off" if os.environ.get("SS4K_NO_PDL") else "on")},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": plan.in_bytes, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "timed_region_s": ms / 1000,
            "hbm_kernels": {"peak_GB/s": peak_hbm, "kernels": hbm_kernels},
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline("cfg2")
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2"])
    ap.add_argument("--clip", type=int, default=48, help="cfg3: owned frames per GPU per step")
    ap.add_argument("--bsvd", default="split", choices=["split", "f16", "auto"],
                    help="cfg3: split = reference constructor init in fp16 hi/lo split precision (parity configuration); "
                         "f16 = trained-like weights (constructor init x 0.5), single-MMA fp16")
    ap.add_argument("--batch", type=int, default=1, help="cfg2: frames per step")
    ap.add_argument("--dtype", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.steps is None:
        args.steps = 8 if args.workload == "cfg3" else 100
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
    if args.workload == "cfg3":
        run_native_cfg3(args, rank, world, local_rank)
    else:
        run_native_cfg2(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
