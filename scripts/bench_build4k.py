"""Per-kernel CUDA-event timing of the config-4 build (4K scene, 5000 lines, depth 30, side 5760): python scripts/bench_build4k.py [reps]"""
import json
import sys
sys.path.insert(0, ".")
import openfdcm_b200 as fdcm
from tests.util import synth_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fm = fdcm.build_cuda_featuremap(synth_scene(3840, 2160, 5000, seed=4000), fdcm.Dt3CudaParameters(30, 5.0, 1.5, fdcm.distance.L2))
fm.rerun()
fdcm.profile(True, reset=True)
for _ in range(n):
    fm.rerun()
rep = fdcm.profile_report()
fdcm.profile(False)
out = {k: round(v["total_ms"] / max(1, v["launches"]), 4) for k, v in rep.items()}
out["total"] = round(sum(out.values()), 4)
print(json.dumps(out))
