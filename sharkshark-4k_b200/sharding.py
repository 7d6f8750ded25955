"""Multi-GPU sharding of the hot path: one process per GPU, no collective inside the nets.

SURVEY.md section 8e: the upscalers are stateless per frame -> frames shard freely; the BSVD denoiser has a
temporal receptive field of exactly +-16 frames (16 BiBufferConvs, src/upscale/model/bsvd/model.py:582-588), so a
contiguous chunk plus a 16-frame halo on each side reproduces the single-stream result.  The only exchange step is
the gather of the finished uint8 frames to the encoder rank (the reference's streamer process,
src/stream/streamer.py:66-153), in stream order.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

BSVD_HALO = 16


@dataclass
class Chunk:
    lo: int        # first owned frame
    hi: int        # one past the last owned frame
    load_lo: int   # first frame that has to be decoded / denoised (owned range + temporal halo, clipped to the clip)
    load_hi: int

    @property
    def owned(self):
        """slice of the locally denoised clip that holds the owned frames"""
        return slice(self.lo - self.load_lo, self.hi - self.load_lo)


def frame_shards(n_frames, world):
    """Contiguous, balanced frame ranges [lo, hi) per rank (sizes differ by at most one)."""
    base, rem = divmod(n_frames, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def bsvd_chunks(n_frames, world, halo=BSVD_HALO):
    """Per-rank chunk with the temporal halo the denoiser needs on each side."""
    return [Chunk(lo, hi, max(0, lo - halo), min(n_frames, hi + halo)) for lo, hi in frame_shards(n_frames, world)]


def gather_frames(local, n_frames, dst=0, group=None):
    """Gather every rank's finished frames [n_local, ...] to ``dst`` in stream order.  Returns the full clip on
    ``dst`` and None elsewhere.  Ranks may own different frame counts: shards are padded to the largest."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    shards = frame_shards(n_frames, world)
    biggest = max(hi - lo for lo, hi in shards)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(shards)], dim=0)
