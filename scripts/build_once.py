"""Build the config-2 map (1080p, depth 30) a few times — target command for ncu captures."""
import sys
sys.path.insert(0, ".")
import openfdcm_b200 as fdcm
from tests.util import synth_scene
d = {"L2": fdcm.distance.L2, "L2_SQUARED": fdcm.distance.L2_SQUARED, "L1": fdcm.distance.L1}[sys.argv[1] if len(sys.argv) > 1 else "L2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fm = fdcm.build_cuda_featuremap(synth_scene(1920, 1080, 2000, seed=2000), fdcm.Dt3CudaParameters(30, 5.0, 1.5, d))
for _ in range(n):
    fm.rerun()
print("ok", fm.width, fm.height)
