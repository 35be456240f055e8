"""ctypes front-end of the CPU oracle (oracle/fdcm_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under openfdcm_b200/ may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfdcm_oracle.so")

L2, L2_SQUARED, L1 = 0, 1, 2

MATCH_DTYPE = np.dtype([("tmpl_idx", "<i4"), ("score", "<f4"), ("transform", "<f4", (6,))])


def build(force=False):
    src = os.path.join(_HERE, "fdcm_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_dt3_build.restype = C.c_void_p
        _lib.orc_search.restype = C.c_long
        _lib.orc_eigen_sum.restype = C.c_float
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def lines_to_records(lines):
    """(4,N) LineArray (reference convention: one column per line, rows x1,y1,x2,y2) -> (N,4) float32
    records (the column-major memory image of the reference's Eigen LineArray)."""
    a = np.asarray(lines, dtype=np.float32)
    if a.ndim == 1 and a.size == 4:
        a = a.reshape(4, 1)
    if a.size == 0:
        return np.zeros((0, 4), np.float32)
    if a.ndim != 2 or a.shape[0] != 4:
        raise ValueError(f"expected a (4,N) line array, got shape {a.shape}")
    return np.ascontiguousarray(a.T, dtype=np.float32)


def records_to_lines(rec):
    return np.ascontiguousarray(np.asarray(rec, dtype=np.float32).reshape(-1, 4).T)


def pack_templates(templates):
    recs = [lines_to_records(t) for t in templates]
    off = np.zeros(len(recs) + 1, dtype=np.int32)
    for i, r in enumerate(recs):
        off[i + 1] = off[i] + r.shape[0]
    flat = np.concatenate(recs, axis=0) if recs else np.zeros((0, 4), np.float32)
    return _f32(flat).reshape(-1, 4), off


# ---- primitives ------------------------------------------------------------------------------
def rasterize_vector(v):
    v = _f32(v)
    out = np.zeros(2, np.float32)
    lib().orc_rasterize_vector(_p(v), _p(out))
    return out


def rasterize_line(line):
    line = _f32(line).reshape(4)
    cap = 1 << 16
    out = np.zeros((cap, 2), np.int64)
    n = lib().orc_rasterize_line(_p(line), _p(out), cap)
    return out[:n].copy()   # rows = (x, y)


def clip_lines(lines, box, delete_oob=True):
    r = lines_to_records(lines)
    out = np.zeros_like(r)
    n = lib().orc_clip_lines(_p(r), r.shape[0], C.c_float(box[0]), C.c_float(box[1]), C.c_float(box[2]),
                             C.c_float(box[3]), int(delete_oob), _p(out))
    return records_to_lines(out[:n])


def draw_lines(img, lines, color):
    img = _f32(img).copy()
    r = lines_to_records(lines)
    lib().orc_draw_lines(_p(img), img.shape[0], img.shape[1], _p(r), r.shape[0], C.c_float(color))
    return img


def distance_transform(lines, size_wh, distance=L2):
    r = lines_to_records(lines)
    W, H = int(size_wh[0]), int(size_wh[1])
    out = np.zeros((H, W), np.float32)
    lib().orc_distance_transform(_p(r), r.shape[0], W, H, int(distance), _p(out))
    return out


def dt_pass_l2_1d(f):
    f = _f32(f).copy()
    lib().orc_dt_pass_l2_1d(_p(f), f.size)
    return f


def line_integral(img, angle):
    img = _f32(img).copy()
    lib().orc_line_integral(_p(img), img.shape[0], img.shape[1], C.c_float(angle))
    return img


def scene_shift(scene, padding):
    r = lines_to_records(scene)
    shift = np.zeros(2, np.float32)
    size = np.zeros(2, np.uint64)
    lib().orc_scene_shift(_p(r), r.shape[0], C.c_float(padding), _p(shift), _p(size))
    return shift, size


def closest_orientation(keys, line):
    keys = _f32(keys)
    line = _f32(line).reshape(4)
    return lib().orc_closest_orientation(_p(keys), keys.size, _p(line))


def angle_keys(depth):
    k = np.zeros(depth, np.float32)
    lib().orc_angle_keys(depth, _p(k))
    return k


def propagate_orientation(planes, keys, coeff):
    planes = _f32(planes).copy()
    keys = _f32(keys)
    D, H, W = planes.shape
    lib().orc_propagate_orientation(_p(planes), D, H, W, _p(keys), C.c_float(coeff))
    return planes


def minmax_translation(tmpl, vec, size_wh, extra=(0.0, 0.0)):
    r = lines_to_records(tmpl)
    vec = _f32(vec)
    extra = _f32(extra)
    out = np.zeros(2, np.float32)
    lib().orc_minmax_translation(_p(r), r.shape[0], _p(vec), C.c_uint64(int(size_wh[0])), C.c_uint64(int(size_wh[1])),
                                 _p(extra), _p(out))
    return out


def eigen_sum(c):
    c = _f32(c)
    return float(lib().orc_eigen_sum(_p(c), c.size))


def align(tmpl_line, scene_line):
    a, b = _f32(tmpl_line).reshape(4), _f32(scene_line).reshape(4)
    t1, t2 = np.zeros(6, np.float32), np.zeros(6, np.float32)
    lib().orc_align(_p(a), _p(b), _p(t1), _p(t2))
    return t1.reshape(2, 3), t2.reshape(2, 3)


def transform(lines, t23):
    r = lines_to_records(lines)
    t = _f32(t23).reshape(6)
    out = np.zeros_like(r)
    lib().orc_transform(_p(r), r.shape[0], _p(t), _p(out))
    return records_to_lines(out)


def default_search(tmpl, scene, max_tmpl_lines, max_scene_lines):
    t, s = lines_to_records(tmpl), lines_to_records(scene)
    cap = int(max_tmpl_lines * max_scene_lines) + 1
    out = np.zeros((cap, 2), np.int64)
    n = lib().orc_default_search(_p(t), t.shape[0], _p(s), s.shape[0], C.c_uint64(max_tmpl_lines),
                                 C.c_uint64(max_scene_lines), _p(out), cap)
    return out[:n].copy()   # rows = (tmplLineIdx, sceneLineIdx)


def filter_in_range(lines, center, lo, hi):
    r = lines_to_records(lines)
    out = np.zeros(max(1, r.shape[0]), np.int64)
    n = lib().orc_filter_in_range(_p(r), r.shape[0], C.c_float(center[0]), C.c_float(center[1]), C.c_float(lo), C.c_float(hi), _p(out))
    return out[:n].tolist()


def concentric_search(tmpl, scene, max_tmpl_lines, max_scene_lines, center, lo, hi):
    t, s = lines_to_records(tmpl), lines_to_records(scene)
    cap = int(max_tmpl_lines * max_scene_lines) + 1
    out = np.zeros((cap, 2), np.int64)
    n = lib().orc_concentric_search(_p(t), t.shape[0], _p(s), s.shape[0], C.c_uint64(max_tmpl_lines), C.c_uint64(max_scene_lines),
                                    C.c_float(center[0]), C.c_float(center[1]), C.c_float(lo), C.c_float(hi), _p(out), cap)
    return out[:n].copy()


def centered_range(center, n, length):
    b, e = C.c_uint64(0), C.c_uint64(0)
    lib().orc_centered_range(C.c_uint64(center), C.c_uint64(n), C.c_uint64(length), C.byref(b), C.byref(e))
    return b.value, e.value


def template_lengths(templates):
    flat, off = pack_templates(templates)
    out = np.zeros(len(off) - 1, np.float32)
    lib().orc_template_lengths(_p(flat), _p(off), len(off) - 1, _p(out))
    return out


def penalize(kind, tau, matches, lengths):
    m = np.ascontiguousarray(matches, dtype=MATCH_DTYPE).copy()
    lengths = _f32(lengths)
    rc = lib().orc_penalize(int(kind), C.c_float(tau), _p(m), C.c_long(m.size), _p(lengths), C.c_long(lengths.size))
    if rc != 0:
        raise IndexError("In penalize, the size of templatelengths is not consistent with match template indices")
    return m


def sort_matches(matches):
    m = np.ascontiguousarray(matches, dtype=MATCH_DTYPE).copy()
    lib().orc_sort_matches(_p(m), C.c_long(m.size))
    return m


def stats_reset():
    lib().orc_stats_reset()


def stats_get():
    a, b = C.c_longlong(0), C.c_longlong(0)
    lib().orc_stats_get(C.byref(a), C.byref(b))
    return a.value, b.value


def hardware_concurrency():
    return lib().orc_hardware_concurrency()


class Dt3Cpu:
    """Restatement of Dt3Cpu / buildCpuFeaturemap (dt3cpu.h:46-63,174-234)."""

    def __init__(self, scene, depth=30, coeff=5.0, padding=2.2, distance=L2, nthreads=0, stage=0):
        r = lines_to_records(scene)
        if nthreads <= 0:
            nthreads = hardware_concurrency()
        self._h = C.c_void_p(lib().orc_dt3_build(_p(r), r.shape[0], int(depth), C.c_float(coeff), C.c_float(padding),
                                                 int(distance), int(nthreads), int(stage)))
        d, W, H = C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        shift = np.zeros(2, np.float32)
        lib().orc_dt3_info(self._h, C.byref(d), C.byref(W), C.byref(H), _p(shift))
        self.depth, self.W, self.H, self.shift = d.value, W.value, H.value, shift
        self.keys = np.zeros(self.depth, np.float32)
        if self.depth:
            lib().orc_dt3_keys(self._h, _p(self.keys))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_dt3_free(self._h)
            self._h = None

    def plane(self, d):
        out = np.zeros((self.H, self.W), np.float32)
        lib().orc_dt3_plane(self._h, int(d), _p(out))
        return out

    def planes(self):
        return np.stack([self.plane(d) for d in range(self.depth)]) if self.depth else np.zeros((0, 0, 0), np.float32)

    def minmax_translation(self, tmpl, vec):
        r = lines_to_records(tmpl)
        vec = _f32(vec)
        out = np.zeros(2, np.float32)
        lib().orc_dt3_minmax_translation(self._h, _p(r), r.shape[0], _p(vec), _p(out))
        return out

    def evaluate(self, tmpl, translations):
        r = lines_to_records(tmpl)
        t = _f32(translations).reshape(-1, 2)
        out = np.zeros(t.shape[0], np.float32)
        lib().orc_dt3_evaluate(self._h, _p(r), r.shape[0], _p(t), t.shape[0], _p(out))
        return out

    def optimize_one(self, tmpl, align_vec, batch):
        r = lines_to_records(tmpl)
        av = _f32(align_vec)
        out = np.zeros(3, np.float32)
        has = lib().orc_optimize_one(self._h, _p(r), r.shape[0], _p(av), C.c_long(batch), _p(out))
        return (bool(has), float(out[0]), out[1:].copy())

    def search(self, templates, scene, max_tmpl_lines, max_scene_lines, batch=10, nthreads=0, want_hyp=False, concentric=None):
        """DefaultMatch search; batch<=0 selects DefaultOptimize. Returns MATCH_DTYPE array
        in hypothesis order (and optionally the (tmpl, tmplLine, sceneLine, rev) hypothesis list)."""
        flat, off = pack_templates(templates)
        s = lines_to_records(scene)
        if nthreads <= 0:
            nthreads = hardware_concurrency()
        T = len(off) - 1
        cap = max(1, 2 * T * int(max_tmpl_lines) * int(max_scene_lines))
        out = np.zeros(cap, MATCH_DTYPE)
        hyp = np.zeros((cap, 4), np.int32)
        n_hyp = C.c_long(0)
        n = lib().orc_search(self._h, _p(flat), _p(off), T, _p(s), s.shape[0], C.c_uint64(max_tmpl_lines),
                             C.c_uint64(max_scene_lines), C.c_long(batch), int(nthreads), _p(out), C.c_long(cap),
                             _p(hyp) if want_hyp else None, C.c_long(cap), C.byref(n_hyp) if want_hyp else None,
                             _p(_f32(concentric)) if concentric is not None else None)
        if want_hyp:
            return out[:n].copy(), hyp[:n_hyp.value].copy()
        return out[:n].copy()
