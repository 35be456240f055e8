"""Config 5 on one GPU (16 x 1080p scenes, 1000 templates): wall per scene of the pipelined batch and of the sequential
loop next to the per-kernel CUDA-event times: python scripts/bench_config5.py [n_scenes]"""
import json
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import openfdcm_b200 as fdcm
from tests.util import plant_instances, synth_scene, synth_templates
n_sc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
t5 = synth_templates(1000, 40, 1920, seed=5100)
sc5 = [plant_instances(synth_scene(1920, 1080, 2000, seed=5000 + s), t5, 1920, 1080, seed=5200 + s) for s in range(n_sc)]
params = fdcm.Dt3CudaParameters(30, 5.0, 1.5)
set5 = fdcm.TemplateSet(t5)
searcher, optimizer, penalty = fdcm.DefaultSearch(4, 4), fdcm.BatchOptimize(10), fdcm.ExponentialPenalty(1.5)
batch = fdcm.SceneBatch(params)
fm = fdcm.build_cuda_featuremap(sc5[0], params)


def pipelined():
    return batch.search_topk(sc5, set5, searcher, optimizer, penalty, 10)


def sequential():
    out = []
    for s in sc5:
        fm.rebuild(s)
        out.append(fdcm.search_topk(fm, set5, None, searcher, optimizer, penalty, 10))
    return out


out = {}
for name, fn in (("pipelined", pipelined), ("sequential", sequential)):
    fn()
    t0 = time.perf_counter()
    for _ in range(3):
        r = fn()
    out[name + "_ms_per_scene"] = round((time.perf_counter() - t0) / 3 / n_sc * 1e3, 4)
fdcm.profile(True, reset=True)
sequential()
rep = fdcm.profile_report()
fdcm.profile(False)
out["kernels_ms_per_scene"] = {k: round(v["total_ms"] / n_sc, 4) for k, v in rep.items()}
out["kernel_sum_ms_per_scene"] = round(sum(v["total_ms"] for v in rep.values()) / n_sc, 4)
t0 = time.perf_counter()
for s in sc5:
    fm.rebuild(s)
out["rebuild_only_ms_per_scene"] = round((time.perf_counter() - t0) / n_sc * 1e3, 4)
print(json.dumps(out))
