import sys
sys.path.insert(0, ".")
import numpy as np
import openfdcm_b200 as fdcm
from tests.util import plant_instances, synth_scene, synth_templates
tm = synth_templates(8, 12, 320, seed=1)
for (w, h, pad, dist) in ((320, 240, 1.5, fdcm.distance.L2), (333, 217, 1.0, fdcm.distance.L2_SQUARED), (200, 150, 2.2, fdcm.distance.L1)):
    scene = plant_instances(synth_scene(w, h, 60, seed=w), tm, w, h, seed=3)
    fm = fdcm.build_cuda_featuremap(scene, fdcm.Dt3CudaParameters(30, 5.0, pad, dist))
    fm.rerun()
    top = fdcm.search_topk(fm, tm, scene, fdcm.DefaultSearch(3, 4), fdcm.BatchOptimize(10), fdcm.ExponentialPenalty(1.5), k=5)
    print(w, h, fm.width, float(fm.plane(3).sum()), top[0])
