"""Shared fixtures: re-statements of the reference's test utilities and the seeded synthetic
scene / template generators of SURVEY.md §8(d) (shared by oracle and CUDA parity tests)."""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F32 = np.float32


def load_kats():
    with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
        return json.load(f)


def make_rotation(angle):
    """tests/test-utils/include/test-utils/utils.h:38-44 (float sin/cos)."""
    a = F32(angle)
    s, c = F32(math.sin(float(a))), F32(math.cos(float(a)))
    return np.array([[c, -s], [s, c]], dtype=F32)


def logspace(start, end, num):
    """utils.h:54-63: float log10, float step, pow in double, stored as float."""
    ls, le = F32(math.log10(F32(start))), F32(math.log10(F32(end)))
    step = F32((le - ls) / F32(num - 1))
    return np.array([F32(math.pow(10.0, float(F32(ls + F32(i) * step)))) for i in range(num)], dtype=F32)


def create_lines(n, length):
    """utils.h:74-87: n lines from the origin, log-spaced angles in [2pi, 4pi]; returns (4,n)."""
    out = np.zeros((4, n), dtype=F32)
    for i, ang in enumerate(logspace(2 * math.pi, 4 * math.pi, n)):
        r = make_rotation(ang)
        out[2, i] = r[0, 0] * F32(length)
        out[3, i] = r[1, 0] * F32(length)
    return out


def rotate_about(line4, rot, pt):
    """core::rotate(line, rotation, rot_point) (math.h:372-378) in float32, for test inputs."""
    rot = np.asarray(rot, F32)
    pt = np.asarray(pt, F32)
    t = (pt - rot @ pt).astype(F32)
    p = np.asarray(line4, F32).reshape(2, 2)          # rows = points
    q = (p @ rot.T + t).astype(F32)
    return q.reshape(4)


def apply_transform(lines, t23):
    t23 = np.asarray(t23, F32)
    pts = np.asarray(lines, F32).T.reshape(-1, 2)
    out = (pts @ t23[:, :2].T + t23[:, 2]).astype(F32)
    return np.ascontiguousarray(out.reshape(-1, 4).T)


# ---- seeded synthetic workloads ---------------------------------------------
def synth_scene(width, height, n_lines, seed, max_len_frac=0.15, min_len=10.0):
    """Random scene lines inside [0,W-1]x[0,H-1]; two 1-px corner lines pin the extent so the
    padded side is deterministic (= ceil(padding*(W-1)+1)). Returns (4,N) float32."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, width - 1, n_lines)
    cy = rng.uniform(0, height - 1, n_lines)
    ang = rng.uniform(0, math.pi, n_lines)
    ln = rng.uniform(min_len, max_len_frac * width, n_lines)
    dx, dy = 0.5 * ln * np.cos(ang), 0.5 * ln * np.sin(ang)
    x1, y1, x2, y2 = cx - dx, cy - dy, cx + dx, cy + dy
    lines = np.stack([x1, y1, x2, y2])
    lines[[0, 2]] = np.clip(lines[[0, 2]], 0, width - 1)
    lines[[1, 3]] = np.clip(lines[[1, 3]], 0, height - 1)
    corners = np.array([[0, 0, 1, 0], [width - 2, height - 1, width - 1, height - 1]], dtype=np.float64).T
    return np.ascontiguousarray(np.concatenate([lines, corners], axis=1), dtype=F32)


def synth_templates(n_tmpl, n_lines, width, seed, min_len=8.0, max_len=60.0):
    """Random templates: line centres in a box of side 0.1W..0.2W about the origin."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_tmpl):
        side = rng.uniform(0.1 * width, 0.2 * width)
        cx = rng.uniform(-side / 2, side / 2, n_lines)
        cy = rng.uniform(-side / 2, side / 2, n_lines)
        ang = rng.uniform(0, math.pi, n_lines)
        ln = rng.uniform(min_len, max_len, n_lines)
        dx, dy = 0.5 * ln * np.cos(ang), 0.5 * ln * np.sin(ang)
        out.append(np.ascontiguousarray(np.stack([cx - dx, cy - dy, cx + dx, cy + dy]), dtype=F32))
    return out


def plant_instances(scene, templates, width, height, seed, count=4):
    """Append rigidly transformed copies of `count` templates to the scene (true positives)."""
    rng = np.random.default_rng(seed)
    extra = []
    for k in range(min(count, len(templates))):
        t = templates[(k * 7919) % len(templates)]
        th = rng.uniform(-math.pi, math.pi)
        c, s = math.cos(th), math.sin(th)
        tx = rng.uniform(0.3 * width, 0.7 * width)
        ty = rng.uniform(0.3 * height, 0.7 * height)
        extra.append(apply_transform(t, [[c, -s, tx], [s, c, ty]]))
    allx = np.concatenate([scene] + extra, axis=1)
    allx[[0, 2]] = np.clip(allx[[0, 2]], 0, width - 1)
    allx[[1, 3]] = np.clip(allx[[1, 3]], 0, height - 1)
    return np.ascontiguousarray(allx, dtype=F32)
