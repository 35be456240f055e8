"""Pins the CPU oracle against the reference's own known-answer tests.

Every test cites the reference test it re-expresses; vectors live in tests/golden/reference_kats.json.
"""
import math

import numpy as np
import pytest

from oracle import fdcm_oracle as orc
from tests.util import F32, apply_transform, create_lines, load_kats, make_rotation, rotate_about

KATS = load_kats()
PI = math.pi


def _angle(expr):
    eps = F32(PI / 12)
    table = {
        "-pi/2": -F32(PI / 2), "-pi/4": F32(-PI / 4.0), "0": F32(0), "pi/4": F32(PI / 4.0), "pi/2": F32(PI / 2),
        "-pi/4-eps": -F32(PI / 4) - eps, "-pi/4+eps": -F32(PI / 4) + eps,
        "pi/4-eps": F32(PI / 4 - float(eps)), "pi/4+eps": F32(PI / 4) + eps,
        "3pi/4-eps": 3 * F32(PI / 4) - eps, "3pi/4+eps": 3 * F32(PI / 4) + eps,
        "-3pi/4-eps": -3 * F32(PI / 4) - eps, "-3pi/4+eps": -3 * F32(PI / 4) + eps,
    }
    return F32(table[expr])


def _val(v):
    if isinstance(v, str):
        t60 = F32(1.0) / F32(math.sqrt(F32(3.0)))
        return {"inf": math.inf, "nan": math.nan, "tan60": float(t60), "-tan60": -float(t60)}[v]
    return float(v)


# ---- dt3cpu.test.cpp --------------------------------------------------------------------------
@pytest.mark.parametrize("case", KATS["scene_centered_translation"], ids=lambda c: c["cite"][-5:])
def test_scene_centered_translation(case):
    shift, size = orc.scene_shift(np.array(case["scene"], F32), case["padding"])
    assert list(size) == case["size"]
    assert np.allclose(shift, case["shift"], rtol=0, atol=1e-5)
    # dt3cpu.test.cpp:50-54: the shifted bbox centre is the centre of the feature window
    s = np.array(case["scene"], F32)
    pts = s.T.reshape(-1, 2)
    new_center = ((pts.max(0) + shift) + (pts.min(0) + shift)) / 2
    assert np.allclose(new_center, (size.astype(F32) - 1) / 2, rtol=0, atol=1e-5)


@pytest.mark.parametrize("case", KATS["minmax_translation"], ids=lambda c: c["cite"])
def test_minmax_translation(case):
    tmpl = np.zeros((4, 0), F32) if case["tmpl"] is None else np.array(case["tmpl"], F32)
    got = orc.minmax_translation(tmpl, case["vec"], case["size"])
    for g, e in zip(got, case["expect"]):
        e = _val(e)
        if math.isnan(e):
            assert math.isnan(g)
        elif math.isinf(e):
            assert g == e
        else:
            assert abs(g - e) <= 1.1920929e-7 + 1e-10 * max(abs(g), abs(e))   # relativelyEqual


def test_closest_orientation():
    """dt3cpu.test.cpp:230-248 (keys include a wrap-around pair and pi)."""
    keys = np.sort(np.array([F32(-F32(PI / 2) + F32(PI / 100)), F32(-PI / 4.0), F32(0), F32(PI / 4.0),
                             F32(F32(PI / 2) - F32(PI / 100)), F32(PI)], F32))
    for angle in keys:
        rot = make_rotation(angle)
        line = np.concatenate([rot @ np.array([0, 0], F32), rot @ np.array([1, 0], F32)]).astype(F32)
        idx = orc.closest_orientation(keys, line)
        x = math.fmod(float(angle) + PI / 2, PI)     # constrainHalfAngle (math.h:218-224)
        x += PI * (x < 0)
        expect = x - PI / 2
        assert abs(float(keys[idx]) - expect) <= 1.2e-7 + 1e-10 * abs(expect)


def test_classify_lines():
    c = KATS["classify_lines"]
    keys = np.array([_angle(k) for k in c["keys"]], F32)
    lines = np.array(c["lines"], F32)
    got = [[] for _ in keys]
    for i in range(lines.shape[1]):
        got[orc.closest_orientation(keys, lines[:, i])].append(i)
    assert got == c["expect"]


def test_propagate_orientation():
    """dt3cpu.test.cpp:268-295."""
    coeff = 0.5
    W, H = 30, 40
    keys = np.array([-F32(PI / 2), -F32(PI / 4), 0, F32(PI / 4)], F32)
    planes = np.full((4, H, W), np.inf, F32)
    planes[0] = orc.distance_transform(np.array([[0], [0], [0], [39]], F32), (W, H), orc.L2)
    out = orc.propagate_orientation(planes, keys, coeff)
    d1 = out[0][0, 29]
    assert d1 == 29.0
    for k, p in zip(keys, out):
        x = math.fmod(float(keys[0] - k) + PI / 2, PI)
        x += PI * (x < 0)
        dist_angle = abs(x - PI / 2)
        assert abs((d1 + dist_angle * coeff) - p[0, 29]) <= 1e-5


def test_build_featuremap_line_cost_le_1():
    """dt3cpu.test.cpp:296-317."""
    scene = np.array([[0, 0, 0, 0, 1], [0, 0, 0, 1, 1], [0, 1, 1, 1, 1], [1, 1, 0, 0, 0]], F32)
    fm = orc.Dt3Cpu(scene, depth=4, coeff=50.0, padding=1.0)
    for i in range(scene.shape[1]):
        line = scene[:, i]
        f = fm.plane(orc.closest_orientation(fm.keys, line))
        p1 = np.round(line[:2]).astype(int)
        p2 = np.round(line[2:]).astype(int)
        assert abs(f[p2[1], p2[0]] - f[p1[1], p1[0]]) <= 1.0


@pytest.mark.parametrize("case", KATS["dt3_golden_rows"], ids=lambda c: c["cite"])
def test_dt3_golden_rows(case):
    scene = np.array(case["scene"], F32)
    fm = orc.Dt3Cpu(scene, depth=case["depth"], coeff=case["coeff"], padding=case["padding"])
    f = fm.plane(orc.closest_orientation(fm.keys, scene[:, 0]))
    row = f[int(f.shape[0] / 2)]
    assert np.allclose(row, np.array(case["middle_row"], F32), rtol=0, atol=1e-5)


# ---- imgproc.test.cpp -------------------------------------------------------------------------
@pytest.mark.parametrize("case", KATS["rasterize_line"], ids=lambda c: c["angle"])
def test_rasterize_line(case):
    ang = _angle(case["angle"])
    line = np.array(case["base"], F32)
    if case["angle"] != "0":
        line = rotate_about(line, make_rotation(ang), case["rot_point"])
    pts = orc.rasterize_line(line)
    assert pts[:, 0].tolist() == case["x"]
    assert pts[:, 1].tolist() == case["y"]


def test_rasterize_short_line():
    """imgproc.test.cpp:86-95."""
    pts = orc.rasterize_line([0, 0, 0.4, 0])
    assert pts.shape[0] == 1 and pts[0].tolist() == [0, 0]


@pytest.mark.parametrize("case", KATS["draw_lines"], ids=lambda c: c["cite"])
def test_draw_lines(case):
    img = np.zeros(case["shape"], F32)
    out = orc.draw_lines(img, np.array(case["lines"], F32), 1.0)
    if case.get("expect_nonzero"):
        assert np.any(out != 0)
    else:
        assert np.array_equal(out, np.array(case["expect"], F32))


def test_draw_lines_empty():
    """imgproc.test.cpp:113-120."""
    img = np.zeros((2, 2), F32)
    assert np.array_equal(orc.draw_lines(img, np.zeros((4, 0), F32), 1.0), img)


def test_line_integral_orientations():
    """imgproc.test.cpp:146-164."""
    base = np.array([8, 8, 11, 8], F32)
    for ang in [-F32(PI / 2), -F32(PI / 4), F32(0), F32(PI / 4), F32(F32(PI / 2) - F32(1e-4))]:
        line = rotate_about(base, make_rotation(ang), [8, 8])
        img = orc.draw_lines(np.zeros((20, 20), F32), line.reshape(4, 1), 1.0)
        out = orc.line_integral(img, ang)
        assert out.max() in (3.0, 4.0)


@pytest.mark.parametrize("name", ["L2", "L1", "L2_SQUARED"])
def test_distance_transform_goldens(name):
    g = KATS["distance_transform"][name]
    dist = getattr(orc, name)
    # "Test for validity" imgproc.test.cpp:171-180: featuresize (5,10), a column of zeros at x=0
    dt = orc.distance_transform(np.array([[0], [0], [0], [9]], F32), (5, 10), dist)
    assert dt[:, 0].sum() == 0
    for i in range(5):
        assert np.allclose(dt[:, i], float(i) ** (2 if name == "L2_SQUARED" else 1))
    assert abs(dt[:, 1].sum() - dt.shape[0]) <= 1e-5
    line = orc.distance_transform(np.array([[2], [0], [5], [0]], F32), (8, 2), dist)
    assert np.allclose(line[0], g["line"], rtol=0, atol=1e-5)
    pt = orc.distance_transform(np.array([[2], [0], [2], [0]], F32), (4, 1), dist)
    assert np.allclose(pt[0], g["single_point"], rtol=0, atol=1e-5)


def test_dt_inplace_aliasing_quirk():
    """the second loop of imgproc.h:123-128 reads already-overwritten values.
    Not pinned by any reference test; this pins the oracle to the literal source behaviour."""
    big = np.finfo(np.float32).max
    out = orc.dt_pass_l2_1d(np.array([0, 4, big, big], F32))
    assert out.tolist() == [0, 1, 4, 5]          # an exact EDT would give [0, 1, 4, 8]


# ---- drawing.test.cpp -------------------------------------------------------------------------
def test_clip_lines():
    c = KATS["clip_lines"]
    for case in c["cases"]:
        got = orc.clip_lines(np.array(case["line"], F32).reshape(4, 1), c["box"])
        if case["expect"] is None:
            assert got.shape[1] == 0
            keep = orc.clip_lines(np.array(case["line"], F32).reshape(4, 1), c["box"], delete_oob=False)
            assert keep.shape[1] == 1
        else:
            assert got[:, 0].tolist() == [float(v) for v in case["expect"]], case


# ---- math.test.cpp ----------------------------------------------------------------------------
def test_transform():
    c = KATS["transform"]
    got = orc.transform(np.array(c["lines"], F32), c["mat"])
    assert np.allclose(got, np.array(c["expect"], F32), rtol=0, atol=1e-5)


def test_align():
    """math.test.cpp:226-248."""
    lines = np.array([[0, 0, 0, 0], [-4, 0, 0, 0], [0, 2, 8, 0], [0, 0, 8, 16]], F32)
    aline = np.array([-1, -1, 1, 1], F32)
    for T in orc.align(lines[:, 0], aline):
        al = orc.transform(lines, T)
        c_ref = (aline[2:] + aline[:2]) / 2
        c_got = (al[2:, 0] + al[:2, 0]) / 2
        assert np.allclose(c_ref, c_got, rtol=0, atol=1e-5)
        d = al[2:] - al[:2]
        d0 = lines[2:] - lines[:2]
        with np.errstate(divide="ignore", invalid="ignore"):
            ang = np.arctan(d[1] / d[0])
            ang0 = np.arctan(d0[1] / d0[0])
        diff = math.atan(1.0) - ang0[0]
        x = np.fmod(ang0 + diff + PI / 2, PI)
        x += PI * (x < 0)
        expect = x - PI / 2
        ok = np.isfinite(ang) & np.isfinite(ang0)
        assert np.allclose(ang[ok], expect[ok], rtol=0, atol=1e-5)


def test_rasterize_vector():
    c = KATS["rasterize_vector"]
    for case in c["cases"]:
        v = make_rotation(_angle(case["angle"])) @ np.array([2.0, 0.0], F32)
        got = orc.rasterize_vector(v)
        assert np.allclose(got, [_val(e) for e in case["expect"]], rtol=0, atol=1e-5), case
    assert np.isnan(orc.rasterize_vector([0, 0])).any()


# ---- searchstrategy.test.cpp -------------------------------------------------------------------
def test_default_search():
    c = KATS["default_search"]
    got = orc.default_search(np.array(c["tmpl"], F32), np.array(c["scene"], F32), c["max_tmpl_lines"], c["max_scene_lines"])
    assert len(got) == 4
    for pair in got.tolist():
        assert pair in c["allowed"]


def test_centered_range():
    for center, n, length, b, e in KATS["centered_range"]["cases"]:
        assert orc.centered_range(center, n, length) == (b, e)


# ---- {batch,default}optimize.test.cpp -----------------------------------------------------------
@pytest.mark.parametrize("batch", [10, 0], ids=["BatchOptimize(10)", "DefaultOptimize"])
@pytest.mark.parametrize("case", KATS["batch_optimize"], ids=lambda c: c["cite"][:30])
def test_optimize(case, batch):
    tmpl = np.array(case["tmpl"], F32)
    if case["pre_transform"] is not None:
        tmpl = orc.transform(tmpl, case["pre_transform"])
    fm = orc.Dt3Cpu(np.array(case["scene"], F32), depth=case["depth"], coeff=case["coeff"], padding=case["padding"])
    has, score, tr = fm.optimize_one(tmpl, case["align_vec"], batch)
    assert has == case["has_value"]
    if not has:
        return
    assert np.allclose(tr, case["translation"], rtol=0, atol=1e-5)
    if "score_exact" in case:
        assert score == case["score_exact"]
    else:
        assert abs(score - case["score_rel"]) <= 1.2e-5 * abs(case["score_rel"])   # Catch WithinRel default


# ---- matchstrategy.test.cpp / tests/python/test_matching.py -------------------------------------
@pytest.mark.parametrize("scene_ratio", [1.0, 0.3])
@pytest.mark.parametrize("nthreads", [1, 2])
def test_default_match_end_to_end(scene_ratio, nthreads):
    """matchstrategy.test.cpp:36-122 (DefaultSearch(3,3), DefaultOptimize, depth 30, padding 2.2)."""
    maxT, maxS, n_lines, length = 3, 3, 10, 10
    tmpl = create_lines(n_lines, length)
    for T in ([[-1, 0, length], [0, -1, length]], [[1, 0, 0], [0, 1, 0]]):
        T = np.array(T, F32)
        scene = orc.transform(tmpl, T)
        fm = orc.Dt3Cpu(scene, depth=30, coeff=5.0, padding=2.2, nthreads=nthreads)
        m = orc.sort_matches(fm.search([tmpl], scene, maxT, maxS, batch=0, nthreads=nthreads))
        assert len(m) == min(maxT, n_lines) * min(n_lines, maxS) * 2
        best = m[0]["transform"].reshape(2, 3)
        # core::allClose(a, b, rtol, atol=1e-5): |a-b| <= atol + rtol*|b| (math.h:203-209); the reference
        # test passes its tolerance in the rtol slot (matchstrategy.test.cpp:64-65)
        assert np.all(np.abs(T[:, :2] - best[:, :2]) <= 1e-5 + 1e-5 * np.abs(best[:, :2]))
        assert np.all(np.abs(T[:, 2] - best[:, 2]) <= 1e-5 + (1.0 / scene_ratio) * np.abs(best[:, 2]))
        assert m[0]["tmpl_idx"] == 0
    # empty scene / no templates / empty template -> no matches
    fm = orc.Dt3Cpu(np.zeros((4, 0), F32), depth=30, coeff=5.0, padding=2.2)
    assert len(fm.search([tmpl], np.zeros((4, 0), F32), maxT, maxS, batch=0)) == 0
    fm = orc.Dt3Cpu(tmpl, depth=30, coeff=5.0, padding=2.2)
    assert len(fm.search([], tmpl, maxT, maxS, batch=0)) == 0
    assert len(fm.search([np.zeros((4, 0), F32)], tmpl, maxT, maxS, batch=0)) == 0


def test_python_end_to_end():
    """tests/python/test_matching.py:45-109 (DefaultSearch(4,10), line length 100, float64 inputs cast to
    f32 at the binding). Mirrors the reference's control flow literally: the 180-degree scene is only used
    with L2 (first loop iteration); `scene`/`scene_transform` carry over, so L1 and L2_SQUARED see the
    identity scene twice."""
    maxT, maxS, n_lines, length = 4, 10, 10, 100

    def mkrot(a):
        return np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])

    tmpl = np.zeros((4, n_lines))
    for i, a in enumerate(np.logspace(np.log10(2 * np.pi), np.log10(4 * np.pi), n_lines)):
        e = mkrot(a) @ np.array([length, 0])
        tmpl[:, i] = [0, 0, e[0], e[1]]

    def app(lines, t):
        return (np.matmul(t[:2, :2], lines.reshape(2, -1)) + t[:2, 2:3]).reshape(4, -1)

    T = np.array([[-1, 0, length], [0, -1, length]], np.float64)
    scene = app(tmpl, T)
    for d in (orc.L2, orc.L1, orc.L2_SQUARED):
        fm = orc.Dt3Cpu(scene, depth=30, coeff=5.0, padding=2.2, distance=d, nthreads=4)
        m = orc.sort_matches(fm.search([tmpl], scene, maxT, maxS, batch=0, nthreads=4))
        assert len(m) == min(maxT, n_lines) * min(n_lines, maxS) * 2
        best = m[0]["transform"].reshape(2, 3)
        assert np.allclose(T[:, :2], best[:, :2], atol=1e-5)
        assert np.allclose(T[:, 2], best[:, 2], atol=1.0)
        T = np.array([[1, 0, 0], [0, 1, 0]], np.float64)
        scene = app(tmpl, T)
        fm = orc.Dt3Cpu(scene, depth=30, coeff=5.0, padding=2.2, distance=d, nthreads=4)
        raw = fm.search([tmpl], scene, maxT, maxS, batch=0, nthreads=4)
        m = orc.sort_matches(orc.penalize(1, 1.5, raw, orc.template_lengths([tmpl])))
        assert len(raw) == maxT * maxS * 2
        best = m[0]["transform"].reshape(2, 3)
        assert np.allclose(T[:, :2], best[:, :2], atol=1e-5)
        assert np.allclose(T[:, 2], best[:, 2], atol=1.0)
        empty = np.zeros((4, 0))
        assert len(orc.Dt3Cpu(empty, 30, 5.0, 2.2, d).search([tmpl], empty, maxT, maxS, batch=0)) == 0
        scene = tmpl
        fm = orc.Dt3Cpu(scene, depth=30, coeff=5.0, padding=2.2, distance=d)
        assert len(fm.search([], scene, maxT, maxS, batch=0)) == 0
        assert len(fm.search([np.zeros((4, 0))], scene, maxT, maxS, batch=0)) == 0


# ---- penaltystrategy.test.cpp -------------------------------------------------------------------
def test_penalties():
    """penaltystrategy.test.cpp:35-142."""
    zero = np.zeros((4, 1), F32)
    m = np.zeros(1, orc.MATCH_DTYPE)
    m["score"] = 1
    for kind, tau in ((0, 0.0), (1, 2.0)):
        out = orc.penalize(kind, tau, m, orc.template_lengths([zero]))
        assert not np.isnan(out["score"][0])
    m2 = np.zeros(2, orc.MATCH_DTYPE)
    m2["tmpl_idx"] = [0, 1]
    m2["score"] = 1
    with pytest.raises(IndexError):
        orc.penalize(0, 0.0, m2, np.zeros(0, F32))
    with pytest.raises(IndexError):
        orc.penalize(1, 2.0, m2, np.zeros(0, F32))
    t1 = orc.transform(create_lines(4, 4), [[-1, 0, 0], [0, -1, 0]])
    t2 = orc.transform(create_lines(3, 3), [[1, 0, 1], [0, 1, 2]])
    lengths = orc.template_lengths([t1, t2])
    for i, t in enumerate((t1, t2)):
        d = t[2:] - t[:2]
        assert abs(lengths[i] - np.sqrt((d * d).sum(0)).sum()) < 1e-4
    out = orc.penalize(0, 0.0, m2, lengths)
    assert np.allclose(out["score"], 1.0 / lengths, rtol=1e-6)
    out = orc.penalize(1, 1.45, m2, lengths)
    assert np.allclose(out["score"], 1.0 / np.power(lengths.astype(np.float64), 1.45), rtol=1e-6)
    assert np.array_equal(out["tmpl_idx"], m2["tmpl_idx"])


def test_eigen_sum_order():
    """App. A.14: the packet-4, 2x-unrolled order differs from a left-to-right sum on adversarial data."""
    rng = np.random.default_rng(7)
    for n in (1, 3, 4, 7, 8, 11, 12, 16, 30, 40, 60, 61):
        c = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 6, n)).astype(F32)
        A = [c[j] for j in range(min(4, n))]
        n4, n8 = n // 4 * 4, n // 8 * 8
        if n4 == 0:
            r = c[0]
            for v in c[1:]:
                r = F32(r + v)
        else:
            if n4 > 4:
                B = [c[4 + j] for j in range(4)]
                for i in range(8, n8, 8):
                    for j in range(4):
                        A[j] = F32(A[j] + c[i + j])
                        B[j] = F32(B[j] + c[i + 4 + j])
                A = [F32(a + b) for a, b in zip(A, B)]
                if n4 > n8:
                    A = [F32(A[j] + c[n8 + j]) for j in range(4)]
            r = F32(F32(A[0] + A[2]) + F32(A[1] + A[3]))
            for v in c[n4:]:
                r = F32(r + v)
        assert orc.eigen_sum(c) == float(r)


def test_concentric_range_strategy():
    """searchstrategy.test.cpp:91-160: filterInRange, empty inputs, 0-centred scene."""
    c = KATS["concentric_range"]
    f = c["filter"]
    assert orc.filter_in_range(np.array(f["lines"], F32), f["center"], f["lo"], f["hi"]) == f["expect"]
    z = c["zero_centered"]
    mt, ms, center, lo, hi = z["args"]
    got = orc.concentric_search(np.array(z["tmpl"], F32), np.array(z["scene"], F32), mt, ms, center, lo, hi)
    assert len(got) == 4 and all(p in z["allowed"] for p in got.tolist())
    assert len(orc.concentric_search(np.array(z["tmpl"], F32), np.zeros((4, 0), F32), mt, ms, center, lo, hi)) == 0
    assert len(orc.concentric_search(np.zeros((4, 0), F32), np.array(z["scene"], F32), mt, ms, center, lo, hi)) == 0
