import sys
sys.path.insert(0, ".")
import numpy as np
import openfdcm_b200 as fdcm
from tests.util import synth_scene
w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fm = fdcm.build_cuda_featuremap(synth_scene(w, h, n, seed=11), fdcm.Dt3CudaParameters(30, 5.0, 1.5))
print("ok", fm.width, fm.height, float(fm.plane(3).sum()))
