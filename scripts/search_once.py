"""Config-3 search (5000 templates x 40 lines vs 1080p) a few times — target command for ncu captures."""
import sys
sys.path.insert(0, ".")
import openfdcm_b200 as fdcm
from tests.util import plant_instances, synth_scene, synth_templates
n_tmpl = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tmpls = synth_templates(n_tmpl, 40, 1920, seed=3001)
scene = plant_instances(synth_scene(1920, 1080, 2000, seed=3000), tmpls, 1920, 1080, seed=3002)
fm = fdcm.build_cuda_featuremap(scene, fdcm.Dt3CudaParameters(30, 5.0, 1.5))
ts = fdcm.TemplateSet(tmpls)
for _ in range(n):
    top = fdcm.search_topk(fm, ts, None, fdcm.DefaultSearch(4, 4), fdcm.BatchOptimize(10), fdcm.ExponentialPenalty(1.5), k=10)
print("ok", top[0])
