// cpp_pipeline.cpp — the reference README flow (README.md:46-82) through the C++ host mirror
// (include/openfdcm_b200/openfdcm_cuda.hpp).  Prints the best matches; exit code 0 on success.
// build: g++ -std=c++17 -Iinclude examples/cpp_pipeline.cpp -Lopenfdcm_b200 -lfdcm_b200 -Wl,-rpath,$PWD/openfdcm_b200
#include <cmath>
#include <cstdio>
#include <random>

#include "openfdcm_b200/openfdcm_cuda.hpp"

using namespace openfdcm::cuda;

int main() {
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> ux(0.f, 639.f), uy(0.f, 479.f), ua(0.f, 3.14159265f), ul(10.f, 96.f);
    LineArray scene;
    for (int i = 0; i < 300; ++i) {
        const float cx = ux(rng), cy = uy(rng), a = ua(rng), l = ul(rng) / 2;
        const float p[4] = {cx - l * std::cos(a), cy - l * std::sin(a), cx + l * std::cos(a), cy + l * std::sin(a)};
        for (int k = 0; k < 4; ++k) scene.push_back(std::fmin(std::fmax(p[k], 0.f), (k % 2) ? 479.f : 639.f));
    }
    std::vector<LineArray> templates(20);
    std::uniform_real_distribution<float> uc(-60.f, 60.f), utl(8.f, 60.f);
    for (auto& t : templates)
        for (int i = 0; i < 30; ++i) {
            const float cx = uc(rng), cy = uc(rng), a = ua(rng), l = utl(rng) / 2;
            t.insert(t.end(), {cx - l * std::cos(a), cy - l * std::sin(a), cx + l * std::cos(a), cy + l * std::sin(a)});
        }
    try {
        Dt3CudaParameters params;
        params.depth = 30; params.dt3Coeff = 5.f; params.padding = 1.5f; params.distance = Distance::L2;
        const Dt3Cuda featuremap = buildCudaFeaturemap(scene, params);
        const Size fs = getFeatureSize(featuremap);
        std::printf("feature size %zu x %zu\n", fs.x, fs.y);
        const DefaultSearch searcher{4, 4};
        const BatchOptimize optimizer{10};
        auto matches = search(DefaultMatch{}, searcher, optimizer, featuremap, templates, scene);
        auto penalized = penalize(ExponentialPenalty{1.5f}, matches, getTemplateLengths(templates));
        sortMatches(penalized);
        std::printf("%zu matches; best: tmpl %d score %g\n", penalized.size(), penalized[0].tmplIdx, penalized[0].score);
        const auto top = searchTopK(searcher, optimizer, ExponentialPenalty{1.5f}, featuremap, templates, scene, 5);
        if (top.empty() || top[0].score != penalized[0].score) { std::printf("top-k mismatch\n"); return 2; }
        const auto combos = establishSearchStrategy(searcher, templates[0], scene);
        const auto mm = minmaxTranslation(featuremap, scene, Point2{1.f, 0.f});
        std::printf("combos %zu, minmax (%g, %g)\n", combos.size(), mm[0], mm[1]);
    } catch (const std::exception& e) {
        std::printf("error: %s\n", e.what());
        return 1;
    }
    return 0;
}
