"""CPU model of the band distance-transform kernels (openfdcm_b200/csrc/dt_band_kernels.cu), checked against the oracle's
literal restatement of _distanceTransformColumnPassL2 (core/imgproc.h:91-130).  It pins, without a GPU, the three
arguments the CUDA kernels rest on:
  * the integer stack algorithm (left / mirrored right stack) and the join of two stacks at their single crossing,
  * the chained base values that reproduce the in-place aliasing of the reference's second loop,
  * the candidate pruning of far bands: the survivors of a loose-pop pass over the row next to a plane's edge rows are a
    superset of the owners of every farther row.
Pure Python on small rows: this is a model of the algorithm, not the product."""
import numpy as np
import pytest

from oracle import fdcm_oracle as orc

BIG = 0xFFFF


class Stack:
    """entries [key = g^2 + v^2, v, bound]; bound = first owned pixel (left stack) or last owned pixel (right stack)"""

    def __init__(self, loose=False):
        self.e = []
        self.slack = 1 if loose else 0

    def take_over(self, v, key, wm1):
        while self.e:
            tk, tv, ts = self.e[-1]
            n, dn = key - tk, 2 * (v - tv)
            if n < (ts - self.slack) * dn:
                self.e.pop()
                continue
            if n >= wm1 * dn:
                return wm1 + 1
            return n // dn + 1
        return 0

    def column(self, v, g, wm1):
        key = g * g + v * v
        start = self.take_over(v, key, wm1)
        if start <= wm1:
            self.e.append([key, v, start])

    def column_rev(self, v, g, wm1):
        key, end = g * g + v * v, wm1
        while self.e:
            tk, tv, ts = self.e[-1]
            n, dn = tk - key, 2 * (tv - v)
            if n >= (ts + self.slack) * dn:
                self.e.pop()
                continue
            if n < 0:
                return
            end = n // dn
            break
        self.e.append([key, v, end])


def joined_envelope(g, xsplit):
    """two stacks + join, as dt_row_band_kernel does it; returns [(key, v, first owned pixel)] in ascending v"""
    n, wm1 = len(g), len(g) - 1
    left, right = Stack(), Stack()
    for v in range(0, xsplit):
        if g[v] != BIG:
            left.column(v, int(g[v]), wm1)
    for v in range(n - 1, xsplit - 1, -1):
        if g[v] != BIG:
            right.column_rev(v, int(g[v]), wm1)
    r = right.e[::-1]
    j, right_start = 0, 0
    while j < len(r):
        key, v, end = r[j]
        j += 1
        start = left.take_over(v, key, wm1)
        if start <= end:
            left.e.append([key, v, start])
            right_start = end + 1
            break
    ent = [tuple(x) for x in left.e]
    for key, v, end in r[j:]:
        ent.append((key, v, right_start))
        right_start = end + 1
    return ent


def fill(ent, n):
    """RowFill: chained bases, then out(q) = base(owner) + (q - v)^2"""
    if not ent:
        return np.full(n, np.finfo(np.float32).max, np.float32)
    base = []
    for k, (key, v, s) in enumerate(ent):
        f = key - v * v
        if s > v:                      # vertex left of its own interval: it contributes the already written out(v)
            j = k - 1
            while ent[j][2] > v:
                j -= 1
            f = base[j] + (v - ent[j][1]) ** 2
        base.append(f)
    out, k = np.zeros(n, np.float64), 0
    for q in range(n):
        while k + 1 < len(ent) and ent[k + 1][2] <= q:
            k += 1
        out[q] = base[k] + (q - ent[k][1]) ** 2
    return out.astype(np.float32)


def literal(g):
    fmax = np.finfo(np.float32).max
    return orc.dt_pass_l2_1d(np.where(g == BIG, fmax, g.astype(np.float64) ** 2).astype(np.float32))


@pytest.mark.parametrize("seed", range(6))
def test_two_stack_join_and_chained_bases_match_the_literal_pass(seed):
    rng = np.random.default_rng(seed)
    for trial in range(250):
        n = int(rng.integers(1, 300))
        g = rng.integers(0, [3, 10, 50, 400, 2800, 40][trial % 6], n)
        if trial % 3 == 0:
            g = np.where(rng.random(n) < 0.7, BIG, g)
        if trial % 7 == 0:
            g = np.where(np.arange(n) < n // 3, g, BIG)
        if trial % 11 == 0:
            g = np.where(np.arange(n) > 2 * n // 3, g, BIG)
        xsplit = int(rng.integers(0, n + 1))
        got = fill(joined_envelope(g, xsplit), n)
        assert np.array_equal(got, literal(g)), (seed, trial, n, xsplit)


def _owners(g):
    """columns that own at least one pixel of the row (strict single left stack)"""
    st = Stack()
    for v in range(len(g)):
        if g[v] != BIG:
            st.column(v, int(g[v]), len(g) - 1)
    own = set()
    for k, (_, v, s) in enumerate(st.e):
        nxt = st.e[k + 1][2] if k + 1 < len(st.e) else len(g)
        if nxt > s:
            own.add(v)
    return own


@pytest.mark.parametrize("seed", range(8))
def test_loose_pass_next_to_the_edge_rows_covers_every_farther_row(seed):
    rng = np.random.default_rng(100 + seed)
    w, h = int(rng.integers(20, 90)), int(rng.integers(30, 80))
    mask = np.zeros((h, w), bool)
    for _ in range(int(rng.integers(1, 7))):                       # a few short segments in the middle rows
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(h // 3, 2 * h // 3))
        dx, dy = int(rng.integers(-12, 13)), int(rng.integers(-6, 7))
        for t in np.linspace(0, 1, 30):
            x, y = int(round(x0 + t * dx)), int(round(y0 + t * dy))
            if 0 <= x < w and 0 <= y < h:
                mask[y, x] = True
    rows = np.flatnonzero(mask.any(1))
    r_top, r_bot = int(rows[0]), int(rows[-1])
    top = np.where(mask.any(0), mask.argmax(0), -1)                # first edge row per column
    bot = np.where(mask.any(0), h - 1 - mask[::-1].argmax(0), -1)  # last edge row per column
    for side, y_ref, far_rows in ((0, r_top - 1, range(r_top - 1, -1, -1)), (1, r_bot + 1, range(r_bot + 1, h))):
        if y_ref < 0 or y_ref >= h:
            continue
        dist = lambda y: np.where(top >= 0, (top - y) if side == 0 else (y - bot), BIG)
        # candidate pass as the kernel runs it: loose pops, left half ascending + right half descending, no join
        g0, xs = dist(y_ref), w // 2
        left, right = Stack(loose=True), Stack(loose=True)
        for v in range(0, xs):
            if g0[v] != BIG:
                left.column(v, int(g0[v]), w - 1)
        for v in range(w - 1, xs - 1, -1):
            if g0[v] != BIG:
                right.column_rev(v, int(g0[v]), w - 1)
        cand = {e[1] for e in left.e} | {e[1] for e in right.e}
        for y in far_rows:
            g = dist(y)
            assert _owners(g) <= cand, (seed, side, y)
            # and the pruned row gives the same result as the full one
            gp = np.where(np.isin(np.arange(w), list(cand)), g, BIG)
            assert np.array_equal(fill(joined_envelope(gp, xs), w), literal(g)), (seed, side, y)


def test_first_minimum_by_bit_pattern_order():
    """search_warp_kernel picks the first minimum of a batch with a warp min-reduction over the float bit patterns
    (batchoptimize.cpp:60 std::min_element on scores that are sums of absolute values): for non-negative floats the
    unsigned order of the bits is the numeric order, so min over bits + lowest index among equals == the sequential scan
    `if (s < bmin)`; batches that hold a NaN take the sequential path in the kernel."""
    rng = np.random.default_rng(5)
    for _ in range(2000):
        n = int(rng.integers(1, 33))
        s = rng.choice([0.0, 1.5, 3.25, 7.0, np.inf, 1e-30, 2.0 ** -140], n).astype(np.float32)
        s[rng.random(n) < 0.3] = np.float32(rng.uniform(0, 10))
        bmin, barg = s[0], 0
        for j in range(1, n):
            if s[j] < bmin:
                bmin, barg = s[j], j
        bits = s.view(np.uint32)
        kmin = bits.min()
        assert int(np.flatnonzero(bits == kmin)[0]) == barg and np.float32(bmin).view(np.uint32) == kmin


# ---- per-band candidate pruning (round 2): every band prunes its own columns -------------------------------------
class RingStack(Stack):
    """Loose stack that only remembers its last `cap` entries: older ones fall off the bottom and are kept as survivors
    (they can no longer be popped); when every remembered entry has been popped the stack is treated as empty, i.e. the
    next vertex starts at pixel 0.  Both deviations only ever keep MORE vertices than the unbounded loose stack."""

    def __init__(self, cap):
        super().__init__(loose=True)
        self.cap, self.fallen = cap, set()

    def column(self, v, g, wm1):
        super().column(v, g, wm1)
        while len(self.e) > self.cap:
            self.fallen.add(self.e.pop(0)[1])

    def survivors(self):
        return self.fallen | {e[1] for e in self.e}


def _vertical_distance(mask, y):
    """g(v) of row y: distance to the nearest edge pixel of column v (BIG: none), what the band records encode"""
    h, w = mask.shape
    g = np.full(w, BIG, np.int64)
    for v in range(w):
        ys = np.flatnonzero(mask[:, v])
        if ys.size:
            g[v] = np.abs(ys - y).min()
    return g


def band_candidates(mask, r0, band, seg, cap):
    """Columns a band [r0, r0 + band) has to visit: columns with an edge pixel inside the band, plus the survivors of
    loose stack passes over the band's first and last row, run independently on column segments of width `seg` (no
    join: a vertex of the row's envelope is also a vertex of the envelope of its own segment)."""
    h, w = mask.shape
    r1 = min(r0 + band, h) - 1
    cand = set(np.flatnonzero(mask[r0:r1 + 1].any(0)).tolist())
    for y in (r0, r1):
        g = _vertical_distance(mask, y)
        for s0 in range(0, w, seg):
            st = RingStack(cap)
            for v in range(s0, min(s0 + seg, w)):
                if g[v] != BIG:
                    st.column(v, int(g[v]), w - 1)
            cand |= st.survivors()
    return cand


def band_candidates_one_sided(mask, r0, band, seg, cap, edges_below):
    """A band with no edge pixel inside and all edge pixels on one side: only the loose pass over the row next to them
    (the band's last row when they lie below it, its first row when they lie above), on 2x finer segments."""
    h, w = mask.shape
    r1 = min(r0 + band, h) - 1
    g = _vertical_distance(mask, r1 if edges_below else r0)
    cand = set()
    for s0 in range(0, w, seg):
        st = RingStack(cap)
        for v in range(s0, min(s0 + seg, w)):
            if g[v] != BIG:
                st.column(v, int(g[v]), w - 1)
        cand |= st.survivors()
    return cand


@pytest.mark.parametrize("seed", range(12))
def test_one_sided_band_candidates_cover_every_row_of_the_band(seed):
    """Bands above (below) every edge pixel of the plane: the survivors of ONE loose pass over the band's last (first) row
    are a superset of the owners of each of its rows, and the envelope built from them alone fills every row exactly."""
    rng = np.random.default_rng(900 + seed)
    w, h = int(rng.integers(20, 100)), int(rng.integers(40, 90))
    band, seg, cap = int(rng.choice([4, 8, 16])), int(rng.choice([3, 5, 9, 16])), int(rng.choice([1, 2, 4, 8]))
    ylo, yhi = sorted(int(v) for v in rng.integers(h // 4, 3 * h // 4, 2))
    mask = np.zeros((h, w), bool)
    for _ in range(int(rng.integers(1, 9))):                          # edge pixels confined to rows [ylo, yhi]
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(ylo, yhi + 1))
        dx, dy = int(rng.integers(-25, 26)), int(rng.integers(-8, 9))
        for t in np.linspace(0, 1, 40):
            x, y = int(round(x0 + t * dx)), int(round(y0 + t * dy))
            if 0 <= x < w and ylo <= y <= yhi:
                mask[y, x] = True
    if seed % 3 == 0:
        mask[ylo, :] = True                                           # a full horizontal line: every column is a vertex
    if not mask.any():
        mask[ylo, w // 2] = True
    for r0 in range(0, h, band):
        r1 = min(r0 + band, h) - 1
        if r1 < ylo or r0 > yhi:                                      # one-sided band
            cand = band_candidates_one_sided(mask, r0, band, seg, cap, edges_below=r1 < ylo)
            for y in range(r0, r1 + 1):
                g = _vertical_distance(mask, y)
                assert _owners(g) <= cand, (seed, r0, y)
                gp = np.where(np.isin(np.arange(w), list(cand)), g, BIG)
                xs = int(rng.integers(0, w + 1))
                assert np.array_equal(fill(joined_envelope(gp, xs), w), literal(g)), (seed, r0, y)


@pytest.mark.parametrize("seed", range(10))
def test_per_band_candidates_cover_every_row_of_the_band(seed):
    """An edge pixel above (below) a band that owns a pixel of one of its rows is strictly nearest on the open segment
    towards that pixel, which crosses the band's first (last) row: it survives a loose pass over that row."""
    rng = np.random.default_rng(500 + seed)
    w, h = int(rng.integers(20, 100)), int(rng.integers(20, 70))
    band, seg, cap = int(rng.choice([4, 8, 16])), int(rng.choice([5, 9, 16, 33])), int(rng.choice([1, 2, 4, 8]))
    mask = np.zeros((h, w), bool)
    for _ in range(int(rng.integers(1, 9))):
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        dx, dy = int(rng.integers(-25, 26)), int(rng.integers(-12, 13))
        for t in np.linspace(0, 1, 40):
            x, y = int(round(x0 + t * dx)), int(round(y0 + t * dy))
            if 0 <= x < w and 0 <= y < h:
                mask[y, x] = True
    if seed % 3 == 0:
        mask[int(rng.integers(0, h)), :] = True                     # a full horizontal line: every column is a vertex
    for r0 in range(0, h, band):
        cand = band_candidates(mask, r0, band, seg, cap)
        for y in range(r0, min(r0 + band, h)):
            g = _vertical_distance(mask, y)
            assert _owners(g) <= cand, (seed, r0, y)
            gp = np.where(np.isin(np.arange(w), list(cand)), g, BIG)
            xs = int(rng.integers(0, w + 1))
            assert np.array_equal(fill(joined_envelope(gp, xs), w), literal(g)), (seed, r0, y)
