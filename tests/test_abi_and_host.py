"""CPU-side tests of the product library: the C ABI loads and exports every symbol the header declares,
and the host-only entry points (orientation bins via the slope table, DefaultSearch, penalties, sort,
template lengths) agree with the oracle.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import openfdcm_b200 as fdcm
from openfdcm_b200 import _lib
from oracle import fdcm_oracle as orc
from tests.util import F32, create_lines, load_kats, synth_scene, synth_templates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KATS = load_kats()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fdcm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdcm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fdcm_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.fdcm_abi_version() == 3


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Dt3Params) == 16
    assert C.sizeof(_lib.SearchParams) == 48
    assert C.sizeof(_lib.Dt3Info) == 40
    assert C.sizeof(_lib.SearchStats) == 32
    assert _lib.MATCH_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(fdcm.FdcmError) as e:
        fdcm.build_cuda_featuremap(np.array([[0], [0], [5], [5]], F32))
    assert e.value.status == _lib.FDCM_ERR_CUDA


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "openfdcm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "fdcm_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("depth", [1, 2, 4, 7, 30, 64])
def test_slope_table_reproduces_atanf_bins(depth):
    rng = np.random.default_rng(depth)
    n = 200000
    ang = rng.uniform(-np.pi, np.pi, n)
    ln = 10.0 ** rng.uniform(-3, 3, n)
    lines = np.zeros((n, 4), F32)
    lines[:, 2] = ln * np.cos(ang)
    lines[:, 3] = ln * np.sin(ang)
    # slopes at and next to every bin boundary, plus the special values
    keys = orc.angle_keys(depth).astype(np.float64)
    mids = np.concatenate([(keys[:-1] + keys[1:]) / 2, [keys[-1] + np.pi / depth / 2, -np.pi / 2, np.pi / 2]])
    t = np.tan(mids).astype(F32)
    near = np.concatenate([np.nextafter(t, F32(np.inf)), t, np.nextafter(t, F32(-np.inf))])
    for k in range(1, 40):
        near = np.concatenate([near, np.nextafter(near[-3 * len(t):], F32(np.inf))])
    m = len(near)
    lines[:m, 0] = 0; lines[:m, 1] = 0; lines[:m, 2] = 1; lines[:m, 3] = near
    lines[m:m + 6] = [[0, 0, 0, 1], [0, 0, 0, -1], [0, 0, 0, 0], [0, 0, 1, 0], [0, 0, -1, 0], [0, 0, -0.0, 1]]
    a, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
    _lib.check(_lib.lib().fdcm_orientation_bins(depth, _lib.ptr(lines), n, _lib.ptr(a), _lib.ptr(b)))
    assert np.array_equal(a, b)
    want = np.array([orc.closest_orientation(orc.angle_keys(depth), lines[i]) for i in range(0, n, 97)])
    assert np.array_equal(a[::97], want)


def test_default_search_matches_oracle_including_length_ties():
    c = KATS["default_search"]
    got = fdcm.establish_search_strategy(fdcm.DefaultSearch(c["max_tmpl_lines"], c["max_scene_lines"]),
                                         np.array(c["tmpl"], F32), np.array(c["scene"], F32))
    assert len(got) == 4 and all(p in c["allowed"] for p in got.tolist())
    scene = synth_scene(640, 480, 300, seed=4)
    for tmpl in synth_templates(6, 30, 640, seed=5) + [create_lines(10, 10), create_lines(40, 100)]:   # createLines: all lengths tie
        for mt, ms in ((4, 4), (16, 16), (3, 10), (50, 400)):
            got = fdcm.establish_search_strategy(fdcm.DefaultSearch(mt, ms), tmpl, scene)
            assert np.array_equal(got, orc.default_search(tmpl, scene, mt, ms))
    assert len(fdcm.establish_search_strategy(fdcm.DefaultSearch(4, 4), np.zeros((4, 0), F32), scene)) == 0


def test_penalize_sort_lengths_match_oracle():
    tmpls = synth_templates(7, 23, 640, seed=9) + [np.zeros((4, 1), F32)]
    assert np.array_equal(np.array(fdcm.get_template_lengths(tmpls), F32), orc.template_lengths(tmpls))
    rng = np.random.default_rng(1)
    m = np.zeros(500, fdcm.MATCH_DTYPE)
    m["tmpl_idx"] = rng.integers(0, len(tmpls), 500)
    m["score"] = rng.uniform(0, 100, 500).astype(F32)
    m["score"][::7] = m["score"][0]                      # ties: std::sort order must agree
    m["transform"] = rng.standard_normal((500, 6)).astype(F32)
    lengths = orc.template_lengths(tmpls)
    for pen, kind, tau in ((fdcm.DefaultPenalty(), 0, 0.0), (fdcm.ExponentialPenalty(1.45), 1, 1.45)):
        got = fdcm.penalize(pen, m, lengths)
        want = orc.penalize(kind, tau, m, lengths)
        assert np.array_equal(got, want)
        assert np.array_equal(fdcm.sort_matches(got), orc.sort_matches(want))
    with pytest.raises(IndexError):
        fdcm.penalize(fdcm.ExponentialPenalty(2.0), m, lengths[:2])
    # Match-object flavour of the API (python/src/matching.cpp:266-307)
    objs = [fdcm.Match(int(r["tmpl_idx"]), float(r["score"]), r["transform"]) for r in m[:20]]
    srt = fdcm.sort_matches(fdcm.penalize(fdcm.DefaultPenalty(), objs, lengths))
    assert [o.score for o in srt] == sorted(o.score for o in srt)


def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu():
    """include/openfdcm_b200/openfdcm_cuda.hpp (C++ mirror of the reference's strategy API) compiles against
    the C ABI; without a device the example must fail with an error, never fall back to a CPU path."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "examples", "cpp_pipeline")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cpp_pipeline.cpp"), "-L" + os.path.join(ROOT, "openfdcm_b200"),
                           "-lfdcm_b200", "-Wl,-rpath,$ORIGIN/../openfdcm_b200", "-o", exe])
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "libfdcm_b200" in r.stdout


def test_concentric_search_matches_oracle():
    z = KATS["concentric_range"]["zero_centered"]
    mt, ms, center, lo, hi = z["args"]
    got = fdcm.establish_search_strategy(fdcm.ConcentricRangeStrategy(mt, ms, center, lo, hi), np.array(z["tmpl"], F32), np.array(z["scene"], F32))
    assert len(got) == 4 and all(p in z["allowed"] for p in got.tolist())
    scene = synth_scene(640, 480, 300, seed=14)
    for tmpl in synth_templates(4, 30, 640, seed=15):
        for lo_r, hi_r in ((0.0, 150.0), (100.0, 260.0), (5000.0, 6000.0)):
            s = fdcm.ConcentricRangeStrategy(4, 6, (320, 240), lo_r, hi_r)
            assert np.array_equal(fdcm.establish_search_strategy(s, tmpl, scene), orc.concentric_search(tmpl, scene, 4, 6, (320, 240), lo_r, hi_r))


def test_packio_read_write_roundtrip(tmp_path):
    """tests/python/test_matching.py:95-105 (write/read round trip) + decoding of a real reference asset."""
    lines = create_lines(100, 10)
    path = str(tmp_path / "test_write_array.lines")
    fdcm.write(path, lines)
    assert np.array_equal(fdcm.read(path), lines)
    fdcm.write(path, np.zeros((4, 0)))                       # overwrite + empty
    assert fdcm.read(path).shape == (4, 0)
    with pytest.raises(RuntimeError):
        fdcm.read(str(tmp_path / "missing.scene"))
    real = os.path.join(ROOT, "tests", "golden", "real_assets", "obj_01")
    scene = fdcm.read(os.path.join(real, "scene_0", "camera_0.scene"))
    assert scene.shape == (4, 557) and scene.dtype == np.float32
    assert 190 < scene.min() and scene.max() < 660
    tm = fdcm.read(os.path.join(real, "templates", "template_0.tmpl"))
    assert tm.shape[0] == 4 and 10 <= tm.shape[1] <= 40


def test_every_real_asset_decodes():
    """all 421 templates (12-33 lines) and 40 scenes (388-763 lines) of the reference's demo assets."""
    import glob
    real = os.path.join(ROOT, "tests", "golden", "real_assets")
    tm = glob.glob(os.path.join(real, "obj_0*", "templates", "*.tmpl"))
    sc = glob.glob(os.path.join(real, "obj_0*", "scene_*", "camera_0.scene"))
    assert (len(tm), len(sc)) == (421, 40)
    for p in tm:
        assert 12 <= fdcm.read(p).shape[1] <= 33
    for p in sc:
        a = fdcm.read(p)
        assert 388 <= a.shape[1] <= 763 and a.min() >= 0 and a.max() <= 800
