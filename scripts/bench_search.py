"""Per-kernel CUDA-event timing of the config-3 search (5000 templates x 40 lines vs 1080p): python scripts/bench_search.py [reps]"""
import json
import sys
sys.path.insert(0, ".")
import numpy as np
import openfdcm_b200 as fdcm
from tests.util import plant_instances, synth_scene, synth_templates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tmpls = synth_templates(5000, 40, 1920, seed=3001)
scene = plant_instances(synth_scene(1920, 1080, 2000, seed=3000), tmpls, 1920, 1080, seed=3002)
fm = fdcm.build_cuda_featuremap(scene, fdcm.Dt3CudaParameters(30, 5.0, 1.5))
ts = fdcm.TemplateSet(tmpls)
args = (fm, ts, None, fdcm.DefaultSearch(4, 4), fdcm.BatchOptimize(10), fdcm.ExponentialPenalty(1.5))
for _ in range(3):
    top = fdcm.search_topk(*args, k=10)
fdcm.profile(True, reset=True)
for _ in range(n):
    top = fdcm.search_topk(*args, k=10)
rep = fdcm.profile_report()
fdcm.profile(False)
out = {k: round(v["total_ms"] / max(1, v["launches"]), 4) for k, v in rep.items()}
out["top1"] = [int(top[0]["tmpl_idx"]), float(top[0]["score"])]
out["checksum"] = float(np.sum(top["score"].astype(np.float64)))
print(json.dumps(out))
