"""CPU model of the four-columns-per-lane L1 row transform (openfdcm_b200/csrc/dt_band_kernels.cu: L1Fill) and of the
warp bit-matrix transposition of dt_col_band_kernel, checked against their definitions.

L1 row call of the reference (core/imgproc.h:137-146 applied to the rows, :178-184): out(x) = min_v g(v) + |x - v|.
The kernel evaluates it as min(x + min_{v<=x}(g(v) - v), -x + min_{v>=x}(g(v) + v)): a right-to-left pass leaves the suffix
minimum of g(v) + v per 128-column chunk, the left-to-right pass keeps the prefix minimum of g(v) - v as a carry; inside a
chunk lane l owns columns 4l .. 4l + 3, runs the two scans serially over its four values and only the lane totals go through
the warp scans."""
import numpy as np

BIG = 1 << 28
NONE = 0xFFFF
COLS = 128


def l1_row_model(g, pitch):
    """g: int array of length pitch, NONE = no edge in the column.  Mirrors L1Fill::init + L1Fill::chunk."""
    nch = (pitch + COLS - 1) // COLS
    gp = np.full(nch * COLS, NONE, np.int64)
    gp[:pitch] = g
    S = np.full(nch + 1, BIG, np.int64)
    run = BIG
    for c in range(nch - 1, -1, -1):                      # init(): suffix minima of g(v) + v per chunk
        x = np.arange(c * COLS, (c + 1) * COLS)
        gc = gp[x]
        m = np.where(gc >= NONE, BIG, gc + x).min()
        run = min(run, m)
        S[c] = run
    out = np.zeros(nch * COLS, np.int64)
    carry = BIG
    for c in range(nch):                                   # chunk(): lanes own four consecutive columns
        a = np.zeros((32, 4), np.int64)
        b = np.zeros((32, 4), np.int64)
        x0 = c * COLS + 4 * np.arange(32)
        for i in range(4):
            gi = gp[x0 + i]
            a[:, i] = np.where(gi >= NONE, BIG, gi - (x0 + i))
            b[:, i] = np.where(gi >= NONE, BIG, gi + (x0 + i))
        for i in (1, 2, 3):
            a[:, i] = np.minimum(a[:, i], a[:, i - 1])   # prefix minima inside the lane
        for i in (2, 1, 0):
            b[:, i] = np.minimum(b[:, i], b[:, i + 1])   # suffix minima inside the lane
        pa, pb = a[:, 3].copy(), b[:, 0].copy()           # inclusive warp scans of the lane totals (doubling steps)
        o = 1
        while o < 32:
            ta = np.concatenate([np.full(o, BIG), pa[:-o]])
            tb = np.concatenate([pb[o:], np.full(o, BIG)])
            pa, pb = np.minimum(pa, ta), np.minimum(pb, tb)
            o <<= 1
        xa = np.minimum(np.concatenate([[BIG], pa[:-1]]), carry)        # exclusive: the lanes before, and the chunks before
        xb = np.minimum(np.concatenate([pb[1:], [BIG]]), S[c + 1])      # the lanes after, and the chunks after
        carry = min(carry, pa[31])
        for i in range(4):
            v = np.minimum(np.minimum(a[:, i], xa) + (x0 + i), np.minimum(b[:, i], xb) - (x0 + i))
            out[x0 + i] = np.where(v >= (BIG >> 1), -1, v)
    return out[:pitch]


def l1_row_definition(g):
    v = np.where(g < NONE)[0]
    if v.size == 0:
        return np.full(g.size, -1, np.int64)
    x = np.arange(g.size)[:, None]
    return (g[v][None, :] + np.abs(x - v[None, :])).min(axis=1)


def test_l1_row_model_matches_the_definition():
    rng = np.random.default_rng(7)
    for pitch in (32, 64, 96, 128, 160, 448, 2880):
        for density in (0.0, 0.01, 0.2, 1.0):
            for _ in range(3):
                g = np.where(rng.random(pitch) < density, rng.integers(0, 3000, pitch), NONE).astype(np.int64)
                assert np.array_equal(l1_row_model(g, pitch), l1_row_definition(g)), (pitch, density)


def warp_bit_transpose_model(a):
    """a: 32 row words (bit c = column c) -> 32 column words (bit r = row r); mirrors warp_bit_transpose."""
    a = [int(x) for x in a]
    m, j = 0x0000FFFF, 16
    while j > 0:
        y = [a[l ^ j] for l in range(32)]
        a = [((a[l] & ~m) | ((y[l] >> j) & m)) & 0xFFFFFFFF if l & j else ((a[l] & m) | ((y[l] << j) & ~m)) & 0xFFFFFFFF
             for l in range(32)]
        m = (m ^ (m << (j >> 1))) & 0xFFFFFFFF
        j >>= 1
    return a


def test_warp_bit_transpose_model():
    rng = np.random.default_rng(11)
    for _ in range(20):
        bits = rng.integers(0, 2, (32, 32))
        rows = [sum(int(bits[r, c]) << c for c in range(32)) for r in range(32)]
        cols = warp_bit_transpose_model(rows)
        for c in range(32):
            assert cols[c] == sum(int(bits[r, c]) << r for r in range(32))
