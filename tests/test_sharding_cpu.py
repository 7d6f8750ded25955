"""Host-side multi-GPU logic on CPU (gloo, world_size 2): frame sharding, BSVD chunk + 16-frame halo, ordered
gather to the encoder rank.  The arithmetic here is the CPU oracle (test infrastructure); the product path runs the
same sharding with the native engine on each rank (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ss4k_b200
from ss4k_b200 import sharding
from oracle import bsvd


def test_frame_shards_cover_everything():
    for n in (1, 7, 16, 600):
        for w in (1, 2, 3, 8):
            sh = sharding.frame_shards(n, w)
            assert sh[0][0] == 0 and sh[-1][1] == n and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [hi - lo for lo, hi in sh]
            assert max(sizes) - min(sizes) <= 1


def test_bsvd_chunks_have_16_frame_halo():
    ch = sharding.bsvd_chunks(600, 8)
    assert ch[0].load_lo == 0 and ch[0].load_hi == 75 + 16 and ch[3].load_lo == 225 - 16 and ch[-1].load_hi == 600
    assert all(c.owned.stop - c.owned.start == 75 for c in ch)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sd = bsvd.build_bsvd32(0, weight_scale=0.5)
    clip = torch.rand(n_frames, 4, 8, 8, generator=torch.Generator().manual_seed(3))
    ch = sharding.bsvd_chunks(n_frames, world)[rank]
    local = bsvd.bsvd_clip(sd, clip[ch.load_lo:ch.load_hi])[ch.owned]          # chunk + halo, keep the owned frames
    frames = (local.clamp(0, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    full = sharding.gather_frames(frames, n_frames, dst=0)
    if rank == 0:
        want = (bsvd.bsvd_clip(sd, clip).clamp(0, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1)
        q.put((tuple(full.shape), int((full.int() - want.int()).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_chunked_denoise_and_gather():
    """41 frames over 2 ranks (21 + 20): every rank denoises its chunk with a 16-frame halo; the gathered clip on
    rank 0 must equal the single-stream result frame for frame."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 41, q)) for r in range(2)]
    for p in procs:
        p.start()
    shape, maxdiff = q.get(timeout=280)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shape == (41, 8, 8, 3) and maxdiff <= 1
