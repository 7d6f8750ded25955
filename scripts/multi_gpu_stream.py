"""cfg3 on N GPUs (torchrun): a synthetic NV12 720p stream is sharded into contiguous chunks with the 16-frame BSVD
halo (sharding.bsvd_chunks); every rank denoises chunk + halo with the native BSVD (NV12 decode in its layout kernel), upscales its
OWNED frames with RRDBNet x2, converts to uint8 and the frames are gathered to rank 0 over NCCL in stream order.
Checks (rank 0): the gathered clip equals the single-GPU result of the whole stream (max |diff| in LSB); prints
frames/s of the sharded run (device time, max over ranks)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import ss4k_b200
from ss4k_b200 import _lib as L, realesrgan, bsvd as nb, sharding
from oracle import rrdbnet, bsvd as ob, colour

H, W = 720, 1280
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
T = int(sys.argv[1]) if len(sys.argv) > 1 else 24 * world
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
torch.manual_seed(0)
sr = realesrgan.NativeRRDBNet(rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval().state_dict(), scale=2, num_block=23, device=local)
den = nb.NativeBSVD(ob.build_bsvd32(0, weight_scale=0.5), device=local, act_mode="auto", out_dtype=torch.float16)
# synthetic NV12 stream: smooth moving pattern + noise, made from RGB with the oracle's encoder (same on every rank)
g = np.random.default_rng(1234)
yy, xx = np.mgrid[0:H, 0:W]
frames = np.empty((T, H * W * 3 // 2), dtype=np.uint8)
for t in range(T):
    rgb = np.stack([(xx + 4 * t) % 256, (yy + 2 * t) % 256, (xx + yy) // 8 % 256], axis=-1).astype(np.float32)
    rgb = np.clip(rgb + g.normal(0, 10, rgb.shape), 0, 255).astype(np.uint8)
    frames[t] = colour.rgb_to_nv12(rgb[None])[0]
nv12 = torch.from_numpy(frames).to(dev)


def run_range(load_lo, load_hi, own):
    """NV12 frames [load_lo, load_hi) -> BSVD clip (NV12 decode + noise map in the engine's layout kernel) -> RRDBNet x2
    on the owned frames -> uint8 NHWC."""
    plan_sr = sr._plan(1, H, W, L.FMT_F16_NCHW, L.FMT_U8_NHWC)
    d = den.denoise_frames(nv12[load_lo:load_hi], H, W, 0.075, nv12=True)[own]      # [n_own, 3, H, W] half
    return torch.cat([plan_sr.run(d[i:i + 1].contiguous()) for i in range(d.shape[0])], dim=0)


ch = sharding.bsvd_chunks(T, world)[rank]
out = run_range(ch.load_lo, ch.load_hi, ch.owned)        # warm-up (plans, graphs)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = run_range(ch.load_lo, ch.load_hi, ch.owned)
full = sharding.gather_frames(out, T, dst=0) if world > 1 else out
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ref = run_range(0, T, slice(0, T))                   # the whole stream on one GPU
    diff = (full.int() - ref.int()).abs()
    print(json.dumps({"what": "cfg3: NV12 720p stream -> BSVD (chunk + 16-frame halo per rank) -> RRDBNet x2 -> uint8, NCCL gather to rank 0",
                      "n_gpus": world, "frames": T, "frames_per_rank_incl_halo": ch.load_hi - ch.load_lo,
                      "ms": ms.item(), "frames/s": 1000 * T / ms.item(),
                      "max_abs_diff_vs_single_gpu_lsb": int(diff.max()), "frames_equal": bool((diff == 0).all())}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
