"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the colour stage around the nets: NV12 <-> RGB.

The reference has no implementation of this step: frames cross its process boundaries as rgb24 through ffmpeg
pipes (src/stream/twitch_realtime_handler/twitchgrabber.py:91-102 on the way in,
src/stream/twitch_stream/output_stream.py:115-175 on the way out) and swscale converts inside ffmpeg.  The only
colour matrices in the tree are the BT.601 ones of the EGVSR training code
(src/upscale/model/egvsr/utils/data_utils.py:55-74).  Parity of this stage is therefore UNPINNED by
construction; the definitions below are this repository's own (BT.709 limited range, as HD video uses):

  nv12_to_rgb : float arithmetic, nearest chroma (each 2x2 block shares its UV sample), clamp to [0, 1]
  rgb_to_nv12 : 15-bit fixed point integer arithmetic (bit-exact target for the kernel), chroma = mean of the 2x2 block

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def nv12_to_rgb(nv12, h, w):
    """nv12: uint8 [N, h*w*3/2] (Y plane then interleaved UV plane) -> float32 [N, 3, h, w] in [0, 1]."""
    nv12 = np.asarray(nv12, dtype=np.uint8).reshape(-1, h * w * 3 // 2)
    n = nv12.shape[0]
    y = nv12[:, :h * w].reshape(n, h, w).astype(np.float32)
    uv = nv12[:, h * w:].reshape(n, h // 2, w // 2, 2).astype(np.float32)
    u = np.repeat(np.repeat(uv[..., 0], 2, axis=1), 2, axis=2)
    v = np.repeat(np.repeat(uv[..., 1], 2, axis=1), 2, axis=2)
    yy = (y - 16.0) * np.float32(1.0 / 219.0)
    cb = (u - 128.0) * np.float32(1.0 / 224.0)
    cr = (v - 128.0) * np.float32(1.0 / 224.0)
    r = yy + np.float32(1.5748) * cr
    g = yy - np.float32(0.187324) * cb - np.float32(0.468124) * cr
    b = yy + np.float32(1.8556) * cb
    return np.clip(np.stack([r, g, b], axis=1), 0.0, 1.0).astype(np.float32)


def rgb_to_nv12(rgb):
    """rgb: uint8 [N, h, w, 3] -> uint8 [N, h*w*3/2]; h even, w even (the kernel additionally needs w % 4 == 0)."""
    rgb = np.asarray(rgb, dtype=np.uint8)
    n, h, w, _ = rgb.shape
    p = rgb.astype(np.int32)
    r, g, b = p[..., 0], p[..., 1], p[..., 2]
    y = 16 + ((5983 * r + 20127 * g + 2032 * b + 16384) >> 15)
    s = p.reshape(n, h // 2, 2, w // 2, 2, 3).sum(axis=(2, 4))
    sr, sg, sb = s[..., 0], s[..., 1], s[..., 2]
    u = 128 + ((-3298 * sr - 11094 * sg + 14392 * sb + 65536) >> 17)
    v = 128 + ((14392 * sr - 13073 * sg - 1319 * sb + 65536) >> 17)
    out = np.empty((n, h * w * 3 // 2), dtype=np.uint8)
    out[:, :h * w] = y.reshape(n, -1).astype(np.uint8)
    out[:, h * w:] = np.stack([u, v], axis=-1).reshape(n, -1).astype(np.uint8)
    return out
