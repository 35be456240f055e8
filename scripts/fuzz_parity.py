"""Randomised bit-exact parity sweep of the CUDA build + search against the CPU oracle (run on a GPU box).
usage: python scripts/fuzz_parity.py [n_cases] [seed]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import openfdcm_b200 as fdcm
from oracle import fdcm_oracle as orc
from tests.util import synth_scene, synth_templates, plant_instances

F32 = np.float32
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
DIST = [("L2", fdcm.distance.L2, orc.L2), ("L2_SQUARED", fdcm.distance.L2_SQUARED, orc.L2_SQUARED), ("L1", fdcm.distance.L1, orc.L1)]


def special_scene(kind, w, h):
    if kind == 0:    # one line
        return np.array([[w * 0.2], [h * 0.3], [w * 0.7], [h * 0.35]], F32)
    if kind == 1:    # vertical and horizontal lines only, some on the borders of the extent
        ls = [[0, 0, w - 1, 0], [0, h - 1, w - 1, h - 1], [0, 0, 0, h - 1], [w - 1, 0, w - 1, h - 1], [w / 2, 0, w / 2, h - 1], [0, h / 2, w - 1, h / 2]]
        return np.array(ls, F32).T.copy()
    if kind == 2:    # a dense cluster in one corner + one far line
        s = synth_scene(max(w // 4, 80), max(h // 4, 60), 40, seed=int(rng.integers(1 << 30)), min_len=4.0)
        far = np.array([[w - 30], [h - 20], [w - 5], [h - 3]], F32)
        return np.ascontiguousarray(np.concatenate([s, far], axis=1), F32)
    if kind == 3:    # all lines in the same orientation bin
        n = 25
        cx, cy = rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)
        ln = rng.uniform(10, 0.3 * w, n)
        th = 0.3 + rng.uniform(-0.02, 0.02, n)
        l = np.stack([cx - ln * np.cos(th) / 2, cy - ln * np.sin(th) / 2, cx + ln * np.cos(th) / 2, cy + ln * np.sin(th) / 2])
        l[[0, 2]] = np.clip(l[[0, 2]], 0, w - 1); l[[1, 3]] = np.clip(l[[1, 3]], 0, h - 1)
        return np.ascontiguousarray(l, F32)
    return None


bad = 0
for case in range(n_cases):
    w, h = int(rng.integers(40, 900)), int(rng.integers(40, 700))
    pad = float(rng.choice([1.0, 1.2, 1.5, 2.2]))
    name, gd, od = DIST[case % 3]
    kind = case % 7
    scene = special_scene(kind, w, h)
    if scene is None:
        scene = synth_scene(w, h, int(rng.integers(3, 200)), seed=int(rng.integers(1 << 30)), max_len_frac=float(rng.uniform(0.3, 0.6)), min_len=4.0)
    g = fdcm.build_cuda_featuremap(scene, fdcm.Dt3CudaParameters(30, 5.0, pad, gd))
    c = orc.Dt3Cpu(scene, 30, 5.0, pad, od)
    ok = (g.width, g.height) == (c.W, c.H)
    nd = 0
    for d in range(30):
        gp, cp = g.plane(d), c.plane(d)
        if not np.array_equal(gp, cp):
            nd += int((gp != cp).sum())
    ok = ok and nd == 0
    # search parity on top
    tm = synth_templates(6, 12, max(w, 100), seed=int(rng.integers(1 << 30)))
    got = fdcm.search_all(g, tm, scene, fdcm.DefaultSearch(3, 4), fdcm.BatchOptimize(10))
    want = c.search(tm, scene, 3, 4, batch=10)
    sok = np.array_equal(got["tmpl_idx"], want["tmpl_idx"]) and np.array_equal(got["score"], want["score"], equal_nan=True) and \
        np.array_equal(got["transform"], want["transform"], equal_nan=True)
    print(f"case {case:2d} kind {kind} {name:10s} scene {w}x{h} pad {pad} map {g.width} lines {scene.shape[1]:3d}: planes {'ok' if ok else f'DIFF {nd}'} search {'ok' if sok else 'DIFF'}", flush=True)
    bad += (not ok) + (not sok)
print("FAILURES", bad)
sys.exit(1 if bad else 0)
