"""Model check of the fused residual-dense-block kernel's cross-CTA protocol (csrc/rdb_fused.cu), on the CPU.

The kernel keeps each CTA's band of output rows through the five convs of a block; what a phase reads from the previous
one is guarded by per-epilogue-warp progress counters.  This test re-states the kernel's integer arithmetic (unit ->
owner CTA, sequence position, "positions < V are complete" from the eight counters, the verified-row cache) in Python,
runs all CTAs under random interleavings with the kernel's lazy publication rule (a warp publishes row i when it
reaches its next row, or at once at the end of a phase) and asserts
  * safety: every row a producer loads has really been written (all four quarter-warps stored it) -- never a stale read;
  * liveness: every interleaving terminates (dependencies only point to earlier phases);
  * the owner formula inverts the unit split exactly.
It is a model of the protocol, not of the CUDA code: the GPU parity tests cover the kernel itself."""
import random

import pytest

PHASES = 5
CHUNKS = [1, 1, 1, 1, 2]


def u0(cta, total, G):
    return cta * total // G


class Cta:
    def __init__(self, idx, G, n_img, strips, H):
        self.idx, self.G, self.n_img, self.strips, self.H = idx, G, n_img, strips, H
        self.totals = [c * n_img * strips * H for c in CHUNKS]
        self.rows = []          # sequence of (phase, unit) this CTA computes, in order
        for p in range(PHASES):
            self.rows += [(p, u) for u in range(u0(idx, self.totals[p], G), u0(idx + 1, self.totals[p], G))]
        self.cum = [0]
        for p in range(PHASES):
            self.cum.append(self.cum[-1] + u0(idx + 1, self.totals[p], G) - u0(idx, self.totals[p], G))
        self.loaded = 0         # rows of the sequence whose inputs the producer has fully requested
        self.done = 0           # rows computed and stored (in sequence order)
        self.ctr = [0] * 8      # published rows per epilogue warp (0..3 even positions, 4..7 odd)
        self.pending = [0] * 8
        self.ok = {}            # producer's verified-row cache per (phase, n, strip): (lo, hi)


def decode(unit, n_img, strips, H):
    y = unit % H
    t = unit // H
    strip = t % strips
    t //= strips
    return t // n_img, t % n_img, strip, y   # chunk, n, strip, y


def input_rows(cta, seq_pos):
    """image rows (of the previous phase's tensor) the producer loads for the row at seq_pos, beyond what the band's earlier
    rows already loaded: the kernel loads input rows band by band: r0 = yb-1 .. r1 = ye."""
    p, u = cta.rows[seq_pos]
    _, n, strip, y = decode(u, cta.n_img, cta.strips, cta.H)
    return p, n, strip, [r for r in (y - 1, y, y + 1) if 0 <= r < cta.H]


def try_load(ctas, c, written):
    """producer of CTA c requests the inputs of its next row if the counters allow it"""
    if c.loaded >= len(c.rows):
        return False
    p, n, strip, rows = input_rows(c, c.loaded)
    G, H, strips = c.G, c.H, c.strips
    total0 = c.totals[0]
    if p > 0:
        for r in rows:
            for s2 in (strip - 1, strip, strip + 1):
                if s2 < 0 or s2 >= strips:
                    continue
                lo, hi = c.ok.get((p, n, s2), (0, 0))
                if lo <= r < hi:
                    continue
                base = (n * strips + s2) * H
                u = base + r
                j = ((u + 1) * G - 1) // total0
                uj0, uj1 = u0(j, total0, G), u0(j + 1, total0, G)
                assert uj0 <= u < uj1                                   # the owner formula inverts the split
                qbase = (p - 1) * (uj1 - uj0)
                q = qbase + (u - uj0)
                o = ctas[j]
                V = min(2 * min(o.ctr[0:4]), 2 * min(o.ctr[4:8]) + 1)
                if not q < V:
                    return False                                        # poll again later
                c.ok[(p, n, s2)] = (r, min(uj0 + (V - qbase), uj1, base + H) - base)
        # safety: everything the TMA loads of this row will read has been written
        for r in rows:
            for s2 in (strip - 1, strip, strip + 1):
                if 0 <= s2 < strips:
                    assert (p - 1, n, s2, r) in written, ("stale read", c.idx, p, n, s2, r)
    c.loaded += 1
    return True


def try_compute(c, written):
    """MMA + epilogue of CTA c finish the next row whose inputs are loaded; lazy publication like the kernel"""
    # a row needs its own inputs and (band streaming) is completed by the NEXT input row: modelled by requiring the
    # producer to be one row ahead, except for the last row of the sequence / of a band
    if c.done >= c.loaded:
        return False
    q = c.done
    p, u = c.rows[q]
    chunk, n, strip, y = decode(u, c.n_img, c.strips, c.H)
    par = q & 1
    for quarter in range(4):
        w = par * 4 + quarter
        c.ctr[w] += c.pending[w]       # the warp's previous store is complete: publish it
        c.pending[w] = 0
    if p < 4:
        written.add((p, n, strip, y))
    last_of_phase = q + 2 >= c.cum[p + 1]
    for quarter in range(4):
        w = par * 4 + quarter
        if last_of_phase:
            c.ctr[w] += 1
        else:
            c.pending[w] = 1
    c.done += 1
    return True


@pytest.mark.parametrize("G,n_img,strips,H,seed", [
    (148, 1, 5, 360, 0),     # RRDBNet x2 trunk at 720p
    (148, 1, 5, 360, 1),
    (37, 1, 3, 30, 2),
    (16, 2, 2, 33, 3),
    (7, 1, 1, 29, 4),        # a single strip: only vertical neighbours
    (148, 1, 8, 540, 5),     # 1080p trunk
])
def test_protocol_is_safe_and_live(G, n_img, strips, H, seed):
    rnd = random.Random(seed)
    ctas = [Cta(i, G, n_img, strips, H) for i in range(G)]
    assert all(c.cum[1] >= 2 for c in ctas)
    written = set()
    total_rows = sum(len(c.rows) for c in ctas)
    done_rows, idle = 0, 0
    order = list(range(G))
    while done_rows < total_rows:
        rnd.shuffle(order)
        progressed = False
        for i in order:
            c = ctas[i]
            # random skew: some CTAs run far ahead, others lag
            for _ in range(rnd.choice((0, 1, 1, 2, 5))):
                a = try_load(ctas, c, written)
                b = try_compute(c, written)
                if b:
                    done_rows += 1
                progressed = progressed or a or b
        idle = 0 if progressed else idle + 1
        assert idle < 50, "deadlock: no CTA can make progress"
    for c in ctas:
        assert c.ctr[0:4] == [(len(c.rows) + 1) // 2] * 4 and c.ctr[4:8] == [len(c.rows) // 2] * 4
