"""Measurements for BASELINE.json configs 3-5 (parity for these configs lives in tests/; this script only times):
  cfg3: BSVD-32 clip (8 frames + temporal halo semantics = one clip) followed by RRDBNet x2, 1280x720 NV12 frames
  cfg4: the same at 1920x1080 with RealESRGANer tiles (tile 512, pad 10)
  cfg5: RRDBNet x4 on 1920x1080, tile sweep
Prints one JSON line per measurement (CUDA events, warm, device-resident inputs)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ss4k_b200
from ss4k_b200 import _lib as L
from ss4k_b200 import realesrgan, bsvd as nbsvd
from oracle import rrdbnet, bsvd as obsvd, colour


def ev_time(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
    torch.manual_seed(0)
    if "cfg3" in which or "cfg4" in which:
        net2 = rrdbnet.RRDBNet(3, 3, 2, 64, 23, 32).eval()
        sr = realesrgan.NativeRRDBNet(net2.state_dict(), scale=2, num_block=23, device=0)
        for name, (h, w), tile in (("cfg3", (720, 1280), 0), ("cfg4", (1080, 1920), 512), ("cfg4", (1080, 1920), 492), ("cfg4", (1080, 1920), 0)):
            if name not in which:
                continue
            F = 8
            for mode, sd in (("f16 (trained-like weights x0.5)", obsvd.build_bsvd32(0, weight_scale=0.5)), ("f16 split (constructor init)", obsvd.build_bsvd32(0)))[:1 if (name == "cfg4" and tile != 512) else 2]:
                den = nbsvd.NativeBSVD(sd, device=0, act_mode="auto", out_dtype=torch.float16)
                srm = sr if tile == 0 else realesrgan.NativeRRDBNet(net2.state_dict(), scale=2, num_block=23, device=0, tile=tile, tile_pad=10)
                srm.out_dtype = torch.float16
                x = torch.rand(1, F, 4, h, w, device="cuda")
                x[:, :, 3] = 0.075

                def step():
                    d = den(x)[0]                      # [F,3,h,w] half
                    return [srm(d[i:i + 1]) for i in range(F)]
                t_den = ev_time(lambda: den(x), 3)
                t_all = ev_time(step, 2)
                print(json.dumps({"config": name, "frame": [h, w], "tile": tile, "bsvd_mode": mode, "frames_per_clip": F,
                                  "bsvd_ms_per_frame": t_den / F, "total_ms_per_frame": t_all / F, "frames/s": 1000 * F / t_all}), flush=True)
                del den
    if "live" in which:
        # The one figure the reference publishes (README.md:20): live 720p -> 1440p with the default wiring =
        # FsrcnnUpscalerService.upscale_multi, SRVGGNetCompact-32 x4 (realesr-general-x4v3), batch 4, colour match,
        # bicubic down to 1440p, uint8 in / out: 24 fps on an RTX 4090 (TensorRT fp16).
        from ss4k_b200 import service
        from oracle import srvgg
        net = srvgg.SRVGGNetCompact(3, 3, 64, 32, 4).eval()
        svc = service.FsrcnnUpscalerService(lr_level=3, device=0, denoising=False, model_name='realesr-general-x4v3',
                                            state_dict=net.state_dict(), batch_size=4, denoise_rate=1.0)
        svc.proc_init()
        svc.output_shape = (1440, 2560)
        frames = torch.randint(0, 256, (4, 720, 1280, 3), dtype=torch.uint8, device="cuda")
        t = ev_time(lambda: svc.upscale(frames), 5)
        m = realesrgan.NativeSRVGG(net.state_dict(), num_conv=32, upscale=4, device=0)
        plan = m._plan(4, 720, 1280, L.FMT_U8_NHWC, L.FMT_F16_NCHW)
        tm = ev_time(lambda: plan.run(frames), 5)
        print(json.dumps({"config": "live default (README.md:20: 24 fps on RTX 4090)", "what": "FsrcnnUpscalerService.upscale, SRVGG-32 x4 720p -> 2880p -> colour match -> bicubic 1440p, batch 4, uint8 in/out",
                          "ms_per_batch": t, "frames/s": 4000 / t, "model_only_ms_per_batch": tm, "model_only_frames/s": 4000 / tm,
                          "model_TFLOP/s": 4 * 2.228 / tm}), flush=True)
    if "cfg5" in which:
        net4 = rrdbnet.RRDBNet(3, 3, 4, 64, 23, 32).eval()
        h, w = 1080, 1920
        x = torch.rand(1, 3, h, w, device="cuda")
        # 1004 / 492 / 236: tile + 2 * tile_pad is a multiple of the kernel's 128-pixel strip (no partly filled strip)
        for tile in (0, 1024, 1004, 512, 492, 256, 236):
            m = realesrgan.NativeRRDBNet(net4.state_dict(), scale=4, num_block=23, device=0, tile=tile, tile_pad=10)
            m.out_dtype = torch.float16
            t = ev_time(lambda: m(x), 2)
            tiles = 1 if tile == 0 else -(-h // tile) * -(-w // tile)
            print(json.dumps({"config": "cfg5", "frame": [h, w], "scale": 4, "tile": tile, "tiles": tiles, "ms_per_frame": t, "frames/s": 1000 / t,
                              "useful_TFLOP/s": 74.346 / t}), flush=True)
            del m
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
