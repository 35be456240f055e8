"""The pybind11 module `openfdcm_b200._openfdcm_cuda` (north_star: "modules/python exposes all of it through the existing
pybind11 bindings"): the reference's Python names (modules/python/src/matching.cpp:62-308) over the CUDA types.
CPU: the module imports and carries the surface; GPU: the reference README flow through it, against the oracle."""
import numpy as np
import pytest

from openfdcm_b200 import _openfdcm_cuda as pyfdcm
from tests.util import F32, plant_instances, synth_scene, synth_templates

REFERENCE_NAMES = ["FeatureMap", "ThreadPool", "OptimizeStrategy", "DefaultOptimize", "BatchOptimize", "PenaltyStrategy", "DefaultPenalty",
                   "ExponentialPenalty", "SearchStrategy", "DefaultSearch", "ConcentricRangeStrategy", "MatchStrategy", "DefaultMatch", "Match",
                   "search", "penalize", "get_template_lengths", "sort_matches", "distance"]
CUDA_NAMES = ["Dt3CudaParameters", "build_cuda_featuremap", "Dt3Cuda", "CudaMatch", "search_topk"]


def test_surface():
    for n in REFERENCE_NAMES + CUDA_NAMES:
        assert hasattr(pyfdcm, n), n
    p = pyfdcm.Dt3CudaParameters(depth=30, dt3Coeff=5.0, padding=1.5, distance=pyfdcm.distance.L2_SQUARED)
    assert (p.depth, p.dt3_coeff, p.padding, p.distance, p.device) == (30, 5.0, 1.5, pyfdcm.distance.L2_SQUARED, 0)
    assert pyfdcm.DefaultSearch(4, 7).get_max_scene_lines() == 7 and pyfdcm.BatchOptimize(10, pyfdcm.ThreadPool(4)).get_batch_size() == 10
    assert pyfdcm.BatchOptimize(batch_size=3, num_threads=2).get_batch_size() == 3
    c = pyfdcm.ConcentricRangeStrategy(4, 5, np.array([3.0, 4.0]), 1.0, 9.0)
    assert c.get_center_position().tolist() == [3.0, 4.0] and c.get_high_radius_boundary() == 9.0
    m = pyfdcm.Match(3, 0.5, np.arange(6).reshape(2, 3))
    assert m.tmpl_idx == 3 and m.transform.shape == (2, 3) and m.transform[1, 2] == 5
    # host-only entry points work without a device (same numbers as the ctypes mirror)
    t = synth_templates(3, 9, 200, seed=1)
    import openfdcm_b200 as fdcm
    assert pyfdcm.get_template_lengths(t) == fdcm.get_template_lengths(t)
    ms = [pyfdcm.Match(0, 9.0, np.zeros((2, 3))), pyfdcm.Match(1, 4.0, np.zeros((2, 3)))]
    assert [x.tmpl_idx for x in pyfdcm.sort_matches(ms)] == [1, 0]
    pen = pyfdcm.penalize(pyfdcm.ExponentialPenalty(tau=1.5), ms, pyfdcm.get_template_lengths(t))
    assert len(pen) == 2 and pen[0].score < 9.0
    with pytest.raises(IndexError):
        pyfdcm.penalize(pyfdcm.DefaultPenalty(), ms, [])
    pairs = pyfdcm.establish_search_strategy(pyfdcm.DefaultSearch(2, 3), t[0], synth_scene(100, 80, 12, seed=2))
    assert pairs.shape == (6, 2)


@pytest.mark.gpu
def test_reference_readme_flow_through_pybind():
    """README.md:46-82 with the CUDA names, float64 inputs like the reference's users pass them."""
    import openfdcm_b200 as fdcm
    from oracle import fdcm_oracle as orc
    tmpls = synth_templates(8, 30, 640, seed=6)
    scene = plant_instances(synth_scene(640, 480, 300, seed=5), tmpls, 640, 480, seed=7)
    scene64 = scene.astype(np.float64)
    fm = pyfdcm.build_cuda_featuremap(scene64, pyfdcm.Dt3CudaParameters(depth=30, dt3Coeff=5.0, padding=1.5), pyfdcm.ThreadPool(4))
    c = orc.Dt3Cpu(scene, 30, 5.0, 1.5)
    assert fm.get_feature_size().tolist() == [c.W, c.H] and np.array_equal(fm.get_scene_translation(), c.shift)
    assert np.array_equal(fm.plane(11), c.plane(11))
    matches = pyfdcm.search(pyfdcm.DefaultMatch(), pyfdcm.DefaultSearch(4, 4), pyfdcm.BatchOptimize(10, pyfdcm.ThreadPool(4)), fm,
                            [t.astype(np.float64) for t in tmpls], scene64)
    want = c.search(tmpls, scene, 4, 4, batch=10)
    assert len(matches) == len(want) > 0
    assert np.array_equal(np.array([m.score for m in matches], F32), want["score"])
    assert np.array_equal(np.stack([m.transform.reshape(6) for m in matches]), want["transform"])
    pen = pyfdcm.penalize(pyfdcm.ExponentialPenalty(tau=1.5), matches, pyfdcm.get_template_lengths(tmpls))
    srt = pyfdcm.sort_matches(pen)
    ref = orc.sort_matches(orc.penalize(1, 1.5, want, orc.template_lengths(tmpls)))
    assert np.array_equal(np.array([m.score for m in srt], F32), ref["score"])
    top = pyfdcm.search_topk(fm, tmpls, scene, pyfdcm.DefaultSearch(4, 4), pyfdcm.BatchOptimize(10), pyfdcm.ExponentialPenalty(1.5), 10)
    assert [m.tmpl_idx for m in top] == fdcm.search_topk(fdcm.build_cuda_featuremap(scene, fdcm.Dt3CudaParameters(30, 5.0, 1.5)), tmpls, scene,
                                                         fdcm.DefaultSearch(4, 4), fdcm.BatchOptimize(10), fdcm.ExponentialPenalty(1.5), 10)["tmpl_idx"].tolist()
    # erased feature map + DefaultOptimize + ConcentricRangeStrategy, and the FeatureMap concept entry points
    erased = pyfdcm.FeatureMap(fm)
    m2 = pyfdcm.search(pyfdcm.CudaMatch(), pyfdcm.ConcentricRangeStrategy(4, 5, [320, 240], 0.0, 200.0), pyfdcm.DefaultOptimize(), erased, tmpls, scene)
    w2 = c.search(tmpls, scene, 4, 5, batch=0, concentric=[320, 240, 0.0, 200.0])
    assert np.array_equal(np.array([m.score for m in m2], F32), w2["score"])
    assert pyfdcm.get_feature_size(erased).tolist() == [c.W, c.H]
    placed = (tmpls[0].T.reshape(-1, 2) + np.array([320, 240], F32)).reshape(-1, 4).T.astype(F32)
    assert np.array_equal(pyfdcm.minmax_translation(fm, placed, [1.0, 0.0]), c.minmax_translation(placed, [1, 0]), equal_nan=True)
    tr = np.array([[0, 0], [3, -2], [-7.5, 4]], F32)
    assert np.array_equal(np.array(pyfdcm.evaluate(fm, [placed], [tr])[0], F32), c.evaluate(placed, tr))
    # empty inputs are not errors
    assert pyfdcm.search(pyfdcm.DefaultMatch(), pyfdcm.DefaultSearch(4, 4), pyfdcm.BatchOptimize(10), fm, [], scene) == []
    assert pyfdcm.search(pyfdcm.DefaultMatch(), pyfdcm.DefaultSearch(4, 4), pyfdcm.BatchOptimize(10), fm, tmpls, np.zeros((4, 0))) == []
