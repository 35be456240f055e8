// Compiled and run by tests/test_gpu_parity.py::test_cpp_type_erased_strategies: the reference's type-erased dispatch
// (featuremap.h:57-124, matchstrategy.h:83-140) with the CUDA types, Eigen-free.
//   1. search(MatchStrategy, SearchStrategy, OptimizeStrategy, FeatureMap, ...) with DefaultSearch / BatchOptimize / Dt3Cuda inside
//      takes the fused device path;
//   2. the same call with a USER-DEFINED search strategy (a plain struct + an establishSearchStrategy overload, the
//      reference's extension mechanism) takes the generic composition (host align + erased optimize on the device);
//   both must return the identical match list;
//   3. copies of the erased objects are independent clones that keep working after the original is gone.
#include <cmath>
#include <cstdio>
#include <random>

#include "openfdcm_b200/strategies.hpp"

using namespace openfdcm::cuda;

namespace user {
struct MySearch { size_t t, s; };   // user-defined strategy: same pairs as DefaultSearch, but unknown to the fused path
inline std::vector<SearchCombination> establishSearchStrategy(const MySearch& m, const LineArray& tmpl, const LineArray& scene) {
    return openfdcm::cuda::establishSearchStrategy(DefaultSearch{m.t, m.s}, tmpl, scene);
}
}   // namespace user

static LineArray random_lines(std::mt19937& rng, int n, float w, float h, float lmin, float lmax, float ox, float oy) {
    std::uniform_real_distribution<float> ux(0.f, w), uy(0.f, h), ua(0.f, 3.14159f), ul(lmin, lmax);
    LineArray out;
    for (int i = 0; i < n; ++i) {
        const float cx = ux(rng) + ox, cy = uy(rng) + oy, a = ua(rng), l = ul(rng) / 2;
        out.insert(out.end(), {cx - l * std::cos(a), cy - l * std::sin(a), cx + l * std::cos(a), cy + l * std::sin(a)});
    }
    return out;
}

static bool same(const std::vector<Match>& a, const std::vector<Match>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i].tmplIdx != b[i].tmplIdx || !(a[i].score == b[i].score)) return false;
        for (int k = 0; k < 6; ++k)
            if (!(a[i].transform[(size_t)k] == b[i].transform[(size_t)k])) return false;
    }
    return true;
}

int main() {
    std::mt19937 rng(7);
    LineArray scene = random_lines(rng, 250, 640.f, 480.f, 10.f, 90.f, 0.f, 0.f);
    scene.insert(scene.end(), {0.f, 0.f, 1.f, 0.f, 638.f, 479.f, 639.f, 479.f});
    std::vector<LineArray> templates;
    for (int t = 0; t < 12; ++t) templates.push_back(random_lines(rng, 20, 90.f, 90.f, 8.f, 50.f, -45.f, -45.f));
    templates.push_back(LineArray{});   // an empty template is skipped (defaultmatch.cpp:54)

    std::vector<Match> fused, generic, cloned;
    {
        const Dt3Cuda map = buildCudaFeaturemap(scene, Dt3CudaParameters{30, 5.f, 1.5f, Distance::L2, 0});
        const FeatureMap fm = map;                       // erased
        const MatchStrategy matcher = DefaultMatch{};
        const OptimizeStrategy optimizer = BatchOptimize{10};
        const SearchStrategy known = DefaultSearch{4, 4};
        const SearchStrategy custom = user::MySearch{4, 4};
        if (!fm.dt3cuda() || known.default_search() == nullptr || custom.default_search() != nullptr) return 2;
        fused = search(matcher, known, optimizer, fm, templates, scene);
        generic = search(matcher, custom, optimizer, fm, templates, scene);
        // erased FeatureMap concept entry points
        const Size sz = getFeatureSize(fm);
        if (sz.x == 0 || sz.x != sz.y || sz.x != map.getFeatureSize().x) { std::printf("size %zu %zu\n", sz.x, sz.y); return 3; }
        const auto mm = minmaxTranslation(fm, templates[0], Point2{1.f, 0.f});
        if (!(mm[0] == mm[0])) { /* NaN: template not inside the map at 0 — fine, only the call matters */ }
        // clones outlive the originals
        FeatureMap fm2 = fm;
        SearchStrategy s2 = known;
        OptimizeStrategy o2 = optimizer;
        MatchStrategy m2 = matcher;
        cloned = search(m2, s2, o2, fm2, templates, scene);
        // concentric range through the erased interface: fused and generic agree as well
        const SearchStrategy conc = ConcentricRangeStrategy{4, 5, Point2{320.f, 240.f}, 0.f, 200.f};
        struct Wrap { ConcentricRangeStrategy c; };
        const auto a = search(matcher, conc, optimizer, fm, templates, scene);
        if (a.empty()) return 4;
        // DefaultOptimize + penalty + sort
        const auto d = search(matcher, known, OptimizeStrategy{DefaultOptimize{}}, fm, templates, scene);
        const PenaltyStrategy pen = ExponentialPenalty{1.5f};
        auto p = penalize(pen, d, getTemplateLengths(templates));
        sortMatches(p);
        for (size_t i = 1; i < p.size(); ++i)
            if (p[i].score < p[i - 1].score) return 5;
    }
    if (fused.empty()) { std::printf("no matches\n"); return 6; }
    if (!same(fused, generic)) { std::printf("fused (%zu) != generic (%zu)\n", fused.size(), generic.size()); return 7; }
    if (!same(fused, cloned)) { std::printf("clone mismatch\n"); return 8; }
    std::printf("type-erased dispatch ok: %zu matches, fused == generic == cloned\n", fused.size());
    return 0;
}
