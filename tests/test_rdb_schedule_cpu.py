"""Model check of the fused residual-dense-block kernel's cross-CTA protocol (csrc/rdb_fused.cu), on the CPU.

The kernel cuts the frame into column strips x row bands (same band boundaries in every strip), one CTA per band, and
keeps a CTA on its band through six phases (conv1..4, conv5 as two chunks).  What a phase reads from an earlier one is
guarded by per-epilogue-warp progress counters (a warp announces a row once two newer stores of its own are in flight); every other phase shifts the bands by half a band (the rows that fall
off the top wrap to the bottom of the strip and are taken last) so that nobody has to wait for them.  This test
re-states the kernel's integer arithmetic (band split, shift / wrap / rotation, row -> owner CTA and processing
position, "positions < V are complete" from the eight counters, the verified-row cache) in Python and checks
  * safety under random interleavings: every row a producer loads has really been written -- never a stale read;
  * liveness: every interleaving terminates (dependencies only point to earlier phases);
  * the point of the shift: when all CTAs run at the same pace and a row becomes visible only several row-times after
    it was computed, NO poll has to wait with the shifted schedule, while the un-shifted schedule waits at every phase.
It is a model of the protocol, not of the CUDA code: the GPU parity tests cover the kernel itself."""
import random

import pytest

PHASES = 6
DEP = [-1, 0, 1, 2, 3, 3]


class Geometry:
    def __init__(self, n_img, strips, H, sms=148, shifted=True):
        self.n_img, self.strips, self.H = n_img, strips, H
        self.B = min(sms // (n_img * strips), H // 4)
        self.half = (H // self.B) // 2
        self.shift = [1 if (p & 1) and p < 4 and shifted else 0 for p in range(PHASES)]
        self.G = n_img * strips * self.B

    def place(self, cta):
        sj, bi = divmod(cta, self.B)
        n, strip = divmod(sj, self.strips)
        return n, strip, bi * self.H // self.B, (bi + 1) * self.H // self.B

    def rot(self, p, v0):
        return self.half - v0 if self.shift[p] and v0 < self.half else 0

    def bands(self, cta):
        """the kernel's next_band(): runs of consecutive image rows (phase, yb, ye), in processing order"""
        n, strip, v0, v1 = self.place(cta)
        nrows = v1 - v0
        out = []
        for p in range(PHASES):
            off = self.half if self.shift[p] else 0
            rot = self.rot(p, v0)
            k = 0
            while k < nrows:
                if k < nrows - rot:
                    v, cnt = v0 + rot + k, nrows - rot - k
                else:
                    v, cnt = v0 + (k - (nrows - rot)), nrows - k
                y = v - off
                if y < 0:
                    y += self.H
                out.append((p, y, y + cnt))
                k += cnt
        return out

    def owner(self, dep, n, strip, r):
        """(cta, sequence position, rows that follow r in the owner's order) of image row r in phase dep"""
        H, B = self.H, self.B
        v = r + (self.half if self.shift[dep] else 0)
        if v >= H:
            v -= H
        bi = ((v + 1) * B - 1) // H
        vj0, vj1 = bi * H // B, (bi + 1) * H // B
        assert vj0 <= v < vj1
        nj = vj1 - vj0
        rot = self.rot(dep, vj0)
        pos = v - vj0 - rot
        if pos < 0:
            pos += nj
        piece_end = vj1 if v >= vj0 + rot else vj0 + rot
        return (n * self.strips + strip) * B + bi, dep * nj + pos, piece_end - v, nj


class Cta:
    def __init__(self, geo, idx):
        self.geo, self.idx = geo, idx
        self.n, self.strip, self.v0, self.v1 = geo.place(idx)
        self.nrows = self.v1 - self.v0
        # output-row sequence, input rows in load order (per band r0 = yb-1 .. r1 = ye, clipped to the image) and, per output
        # row, how many loads must have been issued before it can be computed (its input rows y-1, y, y+1)
        self.rows, self.loads, self.need = [], [], []
        for (p, yb, ye) in geo.bands(idx):
            r0, r1 = max(yb - 1, 0), min(ye, geo.H - 1)
            base = len(self.loads)
            self.loads += [(p, r) for r in range(r0, r1 + 1)]
            for y in range(yb, ye):
                self.rows.append((p, y))
                self.need.append(base + (min(y + 1, r1) - r0) + 1)
        self.loaded = 0
        self.done = 0
        self.ctr = [0] * 8
        self.pending = [0] * 8
        self.ok = {}
        self.waits = 0


def poll(geo, ctas, c, p, r, written, visible_ctr=None):
    """the producer's dependency check for input row r of phase p; returns False if it has to wait.  The bands are
    aligned across strips, so the owners in the strip and its two neighbours sit at the same band / position: one poll
    reads all three counter sets and the verified range is the smallest of the three."""
    dep = DEP[p]
    if dep < 0:
        return True
    lo, hi = c.ok.get(p, (0, 0))
    if not lo <= r < hi:
        V, q, follow = None, None, None
        for s2 in (c.strip - 1, c.strip, c.strip + 1):
            if s2 < 0 or s2 >= geo.strips:
                continue
            j, q, follow, nj = geo.owner(dep, c.n, s2, r)
            ctr = visible_ctr(j) if visible_ctr else ctas[j].ctr
            Vd = min(2 * min(ctr[0:4]), 2 * min(ctr[4:8]) + 1)
            V = Vd if V is None else min(V, Vd)
        if not q < V:
            return False
        c.ok[p] = (r, min(r + min(V - q, follow), geo.H))
    for s2 in (c.strip - 1, c.strip, c.strip + 1):
        if 0 <= s2 < geo.strips:
            assert (dep, c.n, s2, r) in written, ("stale read", c.idx, p, s2, r)
    return True


IN_FLIGHT = 2   # kInFlight of the kernel: TMA stores per epilogue warp that may still be in flight


def compute_row(c, written):
    """the epilogue of the next row of CTA c: store it, announce every row of the warp except the IN_FLIGHT newest"""
    q = c.done
    p, y = c.rows[q]
    par = q & 1
    written.add((p, c.n, c.strip, y))
    for quarter in range(4):
        w = par * 4 + quarter
        c.pending[w] += 1                       # issued
        if c.pending[w] > IN_FLIGHT:
            c.ctr[w] += c.pending[w] - IN_FLIGHT
            c.pending[w] = IN_FLIGHT
    c.done += 1


def flush(c):
    """a warp that runs dry (its next accumulator is not ready) waits for its stores in flight and announces them"""
    for w in range(8):
        c.ctr[w] += c.pending[w]
        c.pending[w] = 0


@pytest.mark.parametrize("n_img,strips,H,sms,seed", [
    (1, 5, 360, 148, 0),     # RRDBNet x2 trunk at 720p: 5 strips x 29 bands
    (1, 5, 360, 148, 1),
    (1, 8, 540, 148, 2),     # 1080p trunk: 8 strips x 18 bands
    (2, 2, 33, 16, 3),
    (1, 1, 29, 7, 4),        # a single strip
    (1, 3, 30, 37, 5),       # bands limited by H / 4
])
def test_protocol_is_safe_and_live(n_img, strips, H, sms, seed):
    geo = Geometry(n_img, strips, H, sms)
    rnd = random.Random(seed)
    ctas = [Cta(geo, i) for i in range(geo.G)]
    # every image row of every strip is produced exactly once per phase
    for p in range(PHASES):
        rows = sorted((c.n, c.strip, y) for c in ctas for (pp, y) in c.rows if pp == p)
        assert rows == sorted((n, s, y) for n in range(n_img) for s in range(strips) for y in range(H))
    written = set()
    total_rows = sum(len(c.rows) for c in ctas)
    done_rows, idle = 0, 0
    order = list(range(geo.G))
    while done_rows < total_rows:
        rnd.shuffle(order)
        progressed = False
        for i in order:
            c = ctas[i]
            for _ in range(rnd.choice((0, 1, 1, 2, 5))):     # random skew: some CTAs run ahead, others lag
                if c.loaded < len(c.loads):
                    p, r = c.loads[c.loaded]
                    if c.loaded and p != c.loads[c.loaded - 1][0]:
                        c.ok = {}
                    if poll(geo, ctas, c, p, r, written):
                        c.loaded += 1
                        progressed = True
                if c.done < len(c.rows) and c.loaded >= c.need[c.done]:
                    compute_row(c, written)
                    done_rows += 1
                    progressed = True
                elif any(c.pending):
                    flush(c)
                    progressed = True
        idle = 0 if progressed else idle + 1
        assert idle < 50, "deadlock: no CTA can make progress"
    for c in ctas:   # announced + still in flight == rows of the warp's parity class
        assert [a + b for a, b in zip(c.ctr, c.pending)] == [(len(c.rows) + 1) // 2] * 4 + [len(c.rows) // 2] * 4


def lockstep_makespan(geo, lag):
    """All CTAs compute one row per tick; counter updates become visible `lag` ticks after they were made (on top of the
    kernel's lazy publication: a warp publishes row i when it reaches its next row); the producer requests an input row
    when the row above it is about to be computed.  Returns the ticks until every CTA is done."""
    ctas = [Cta(geo, i) for i in range(geo.G)]
    history = {c.idx: [] for c in ctas}          # (tick, counter snapshot)
    written = set()
    tick = 0
    pos = {c.idx: 0 for c in ctas}               # index into c.loads
    while any(c.done < len(c.rows) for c in ctas):
        tick += 1
        assert tick < 10000

        def visible(j):
            for (t, snap) in reversed(history[j]):
                if t <= tick - lag:
                    return snap
            return [0] * 8

        for c in ctas:
            if c.done >= len(c.rows):
                continue
            stalled = False
            while pos[c.idx] < c.need[c.done]:
                lp, lr = c.loads[pos[c.idx]]
                if pos[c.idx] and lp != c.loads[pos[c.idx] - 1][0]:
                    c.ok = {}
                if not poll(geo, ctas, c, lp, lr, written, visible):
                    stalled = True
                    break
                pos[c.idx] += 1
            if not stalled:
                compute_row(c, written)
            else:
                flush(c)
        for c in ctas:
            history[c.idx].append((tick, list(c.ctr)))
    return tick


def test_half_band_shift_hides_the_publication_latency():
    """720p trunk (5 strips x 29 bands of 12-13 rows); a warp announces a row only when two newer stores of its own are
    in flight (store completion takes a few row-times) and the counter is seen one more row-time later: with the
    half-band shift the six phases take exactly 6 x 13 row-times -- nobody ever waits for data, only for the 13-row
    bands -- while the un-shifted schedule pays the latency at every phase change."""
    ideal = 6 * 13
    shifted = lockstep_makespan(Geometry(1, 5, 360, 148, shifted=True), lag=1)
    plain = lockstep_makespan(Geometry(1, 5, 360, 148, shifted=False), lag=1)
    print(f"row-times for six phases: ideal {ideal}, shifted {shifted}, un-shifted {plain}")
    assert shifted == ideal
    assert plain >= ideal + 5
