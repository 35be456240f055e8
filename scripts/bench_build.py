"""Per-kernel CUDA-event timing of the config-2 build (1080p, depth 30): python scripts/bench_build.py [L2|L2_SQUARED|L1] [reps]"""
import json
import sys
sys.path.insert(0, ".")
import openfdcm_b200 as fdcm
from tests.util import synth_scene
d = {"L2": fdcm.distance.L2, "L2_SQUARED": fdcm.distance.L2_SQUARED, "L1": fdcm.distance.L1}[sys.argv[1] if len(sys.argv) > 1 else "L2"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fm = fdcm.build_cuda_featuremap(synth_scene(1920, 1080, 2000, seed=2000), fdcm.Dt3CudaParameters(30, 5.0, 1.5, d))
for _ in range(3):
    fm.rerun()
fdcm.profile(True, reset=True)
for _ in range(n):
    fm.rerun()
rep = fdcm.profile_report()
fdcm.profile(False)
import time
t0 = time.perf_counter()
for _ in range(n):
    fm.rerun(wait=False)
fm.rerun()
wall = (time.perf_counter() - t0) / (n + 1) * 1e3
out = {k: round(v["total_ms"] / max(1, v["launches"]), 4) for k, v in rep.items()}
out["sum_of_kernels"] = round(sum(out.values()), 4)
out["wall_ms_per_build"] = round(wall, 4)
print(json.dumps(out))
