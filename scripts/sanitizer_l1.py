"""compute-sanitizer target for the L1 kernels: fused L1 build, stage-wise L1 (stand-alone row kernel), a depth without a
fused kernel, odd widths: compute-sanitizer --tool memcheck|racecheck python scripts/sanitizer_l1.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import openfdcm_b200 as fdcm
from tests.util import synth_scene
for (w, h, pad, depth, stage) in ((200, 150, 2.2, 30, 0), (333, 217, 1.0, 30, 1), (130, 97, 1.5, 12, 0), (517, 300, 1.3, 30, 0)):
    scene = synth_scene(w, h, 60, seed=w)
    p = fdcm.Dt3CudaParameters(depth, 5.0, pad, fdcm.distance.L1)
    fm = fdcm.build_cuda_featuremap(scene, p) if stage == 0 else fdcm.build_cuda_featuremap(scene, p, stage=stage)
    print(w, h, fm.width, float(np.nan_to_num(fm.plane(1)).sum()))
