"""Quick timing probe on a GPU box: build at 1080p (config 2) and search (config 3)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import openfdcm_b200 as fdcm
from tests.util import plant_instances, synth_scene, synth_templates

n_tmpl = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
scene = synth_scene(1920, 1080, 2000, seed=2000)
out = {}
fdcm.profile(True, reset=True)
for name, d in (("L2", fdcm.distance.L2), ("L2_SQUARED", fdcm.distance.L2_SQUARED), ("L1", fdcm.distance.L1)):
    t0 = time.time()
    fm = fdcm.build_cuda_featuremap(scene, fdcm.Dt3CudaParameters(30, 5.0, 1.5, d))
    t1 = time.time()
    for _ in range(3):
        fm.rerun()
    fdcm.profile(True, reset=True)
    t2 = time.time()
    for _ in range(5):
        fm.rerun()
    t3 = time.time()
    out[name] = {"first_build_s": t1 - t0, "rerun_ms": (t3 - t2) / 5 * 1e3, "size": [fm.width, fm.height],
                 "kernels": {k: v["total_ms"] / v["launches"] for k, v in fdcm.profile_report().items()}}
    print(name, json.dumps(out[name]), flush=True)
    del fm

tmpls = synth_templates(n_tmpl, 40, 1920, seed=3001)
scene3 = plant_instances(synth_scene(1920, 1080, 2000, seed=3000), tmpls, 1920, 1080, seed=3002)
fm = fdcm.build_cuda_featuremap(scene3, fdcm.Dt3CudaParameters(30, 5.0, 1.5))
t0 = time.time()
ts = fdcm.TemplateSet(tmpls)
t1 = time.time()
s, o, p = fdcm.DefaultSearch(4, 4), fdcm.BatchOptimize(10), fdcm.ExponentialPenalty(1.5)
for _ in range(2):
    top = fdcm.search_topk(fm, ts, scene3, s, o, p, k=10)
fdcm.profile(True, reset=True)
t2 = time.time()
for _ in range(5):
    top = fdcm.search_topk(fm, ts, scene3, s, o, p, k=10)
t3 = time.time()
st = fm.last_search_stats()
out["search"] = {"templates": n_tmpl, "template_upload_s": t1 - t0, "search_ms": (t3 - t2) / 5 * 1e3, "stats": st,
                 "kernels": {k: v["total_ms"] / v["launches"] for k, v in fdcm.profile_report().items()},
                 "top": [[int(r["tmpl_idx"]), float(r["score"])] for r in top]}
print("search", json.dumps(out["search"]), flush=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
