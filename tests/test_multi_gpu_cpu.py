"""Host logic of the multi-GPU front (multi_gpu.MultiGpuUpscalerService: the reference service's interface,
src/upscale/base_service.py:13-110, in front of one worker process per GPU) and of the live-stream chunker, on the
CPU: fake workers stand in for the per-GPU services, so ordering, the drop policy and error propagation are tested
without a device."""
import queue
import random
import time

import pytest
import torch

from ss4k_b200 import multi_gpu, sharding
from ss4k_b200.service import UpscalerQueueEntry


class _FakeService:
    """stands in for FsrcnnUpscalerService in a worker: 'upscales' by repeating pixels; later jobs may finish first"""

    def __init__(self, device, fail_on=None, delay=0.0, **kw):
        self.device, self.fail_on, self.delay = device, fail_on, delay
        self.output_shape = None
        self.kw = kw

    def proc_init(self):
        self.rnd = random.Random(self.device)

    def proc_cleanup(self):
        pass

    def proc_job_recieved(self, job):
        time.sleep(self.delay * self.rnd.random())
        if self.fail_on is not None and job.step == self.fail_on:
            raise RuntimeError(f"worker {self.device} cannot process step {job.step}")
        f = job.frames
        up = f.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        return UpscalerQueueEntry(frames=up, step=job.step, audio_segment=job.audio_segment, elapsed=float(self.device))


def fake_factory(device, **kw):
    return _FakeService(device, **kw)


def _job(step):
    return UpscalerQueueEntry(frames=torch.full((2, 4, 6, 3), step % 251, dtype=torch.uint8), step=step,
                              audio_segment=torch.tensor([step]))


def test_results_come_back_in_arrival_order():
    svc = multi_gpu.MultiGpuUpscalerService(devices=[0, 1, 2], worker_factory=fake_factory, delay=0.02, lr_level=3)
    assert svc.lr_shape == (720, 1280)
    svc.output_shape = (1440, 2560)
    svc.start(ready_timeout=120)
    try:
        n = 40
        for i in range(n):
            svc.push_job(_job(i))
        got = [svc.get_result(timeout=30) for _ in range(n)]
    finally:
        svc.stop()
    assert [e.step for e in got] == list(range(n))
    assert all(tuple(e.frames.shape) == (2, 8, 12, 3) and int(e.frames[0, 0, 0, 0]) == e.step % 251 for e in got)
    assert {int(e.elapsed) for e in got} == {0, 1, 2}          # every worker took part (round-robin)
    assert all(int(e.elapsed) == e.step % 3 for e in got)


def test_on_queue_callback_and_dropped_jobs_do_not_stall_the_reorder_buffer():
    seen = []
    svc = multi_gpu.MultiGpuUpscalerService(devices=[0, 1], worker_factory=fake_factory, on_queue=lambda e: seen.append(e.step),
                                            queue_size=2, delay=0.05)
    svc.start(ready_timeout=120)
    dropped = []
    try:
        for i in range(30):
            try:
                svc.push_job_nowait(_job(i))                    # the reference's frame_skips policy (pipeline.py:76,111)
            except queue.Full:
                dropped.append(i)
            time.sleep(0.005)
        deadline = time.time() + 30
        while len(seen) + len(dropped) < 30 and time.time() < deadline:
            time.sleep(0.01)
    finally:
        svc.stop()
    assert dropped, "the test needs at least one dropped job"
    assert seen == [i for i in range(30) if i not in dropped]   # in order, nothing lost, nothing waited for


def test_worker_exception_is_reported_and_later_jobs_still_flow():
    svc = multi_gpu.MultiGpuUpscalerService(devices=[0, 1], worker_factory=fake_factory, fail_on=3)
    svc.start(ready_timeout=120)
    try:
        for i in range(8):
            svc.push_job(_job(i))
        got, err = [], None
        deadline = time.time() + 30
        while len(got) < 7 and time.time() < deadline:
            try:
                got.append(svc.get_result(timeout=1).step)
            except multi_gpu.WorkerError as e:
                err = str(e)
            except queue.Empty:
                pass
    finally:
        svc.stop()
    assert got == [0, 1, 2, 4, 5, 6, 7]
    assert err is not None and "cannot process step 3" in err


# ---------------------------------------------------------------------------------------------- live-stream chunker
def test_stream_chunker_matches_the_offline_chunks():
    T, world, L = 200, 3, 24
    ch = sharding.StreamChunker(world, L)
    got = []
    fed = 0
    rnd = random.Random(1)
    while fed < T:
        n = min(rnd.choice((1, 4, 7)), T - fed)
        fed += n
        for rank, c in ch.push(n):
            assert c.load_hi <= fed, "a chunk was dispatched before its trailing halo arrived"
            got.append((rank, c))
        assert ch.oldest_needed() <= max(0, ch.next_chunk * L - 16)
    got += ch.finish()
    assert [r for r, _ in got] == [k % world for k in range(len(got))]
    # the same owned ranges / halos as the offline split of the whole clip into ceil(T / L) chunks
    assert [(c.lo, c.hi) for _, c in got] == [(k * L, min((k + 1) * L, T)) for k in range((T + L - 1) // L)]
    for _, c in got:
        assert c.load_lo == max(0, c.lo - 16) and c.load_hi == min(T, c.hi + 16)
        assert c.owned == slice(c.lo - c.load_lo, c.hi - c.load_lo)
    assert sum(c.hi - c.lo for _, c in got) == T
