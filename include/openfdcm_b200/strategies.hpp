// strategies.hpp — the reference's five type-erased "concepts" for the CUDA hot paths, Eigen-free.
//
// Mirrors, name for name, the Concept / Model<T> pimpl + free-function dispatch of openfdcm::matching
// (reference modules/matching/include/openfdcm/matching/):
//   FeatureMap        featuremap.h:57-124       (clone() = deep copy of the Model; a Dt3Cuda model copies a ref-counted handle)
//   SearchStrategy    searchstrategy.h:72-124
//   OptimizeStrategy  optimizestrategy.h:66-118
//   MatchStrategy     matchstrategy.h:83-140
//   PenaltyStrategy   penaltystrategy.h:60-107
// A concrete type T takes part by providing the free functions the reference specialises for its own types
// (getFeatureSize / minmaxTranslation / evaluate, establishSearchStrategy, optimize, search, penalize): the Model<T>
// forwards to them, found by argument-dependent lookup — the same extension mechanism, without the tag bases.
//
// search(DefaultMatch, ...) is the reference's driver (matching/src/matchstrategies/defaultmatch.cpp:32-89) with one
// addition: when the erased arguments hold DefaultSearch / ConcentricRangeStrategy, BatchOptimize / DefaultOptimize and a
// Dt3Cuda, the whole pipeline runs as ONE fused device search (fdcm_search); any other combination (a user-defined
// strategy, another feature map type) takes the generic composition below, which is the reference's loop verbatim:
// establishSearchStrategy -> align -> transform -> optimize(erased) -> combine.
#pragma once
#include <cmath>
#include <memory>
#include <utility>

#include "openfdcm_cuda.hpp"

namespace openfdcm::cuda {

// ---- small core::math pieces the generic driver needs (core/math.h, same evaluation order, no FMA contraction) ----
namespace core {
inline Point2 normalize(const float* l) {                       // math.h:331-333 (Eigen: z > 0 ? v / sqrt(z) : v)
    const float dx = l[2] - l[0], dy = l[3] - l[1];
    const float z = dx * dx + dy * dy;
    if (z > 0.f) {
        const float n = std::sqrt(z);
        return {dx / n, dy / n};
    }
    return {dx, dy};
}
inline std::array<Mat23, 2> align(const float* tl, const float* sl) {   // math.h:387-406
    const Point2 t = normalize(tl), a = normalize(sl);
    const float c = a.x * t.x + a.y * t.y;
    const float s = a.y * t.x - a.x * t.y;
    const float cx = (sl[2] + sl[0]) / 2, cy = (sl[3] + sl[1]) / 2;
    std::array<Mat23, 2> out;
    for (int rev = 0; rev < 2; ++rev) {
        const float r00 = rev ? -c : c, r01 = rev ? s : -s, r10 = rev ? -s : s, r11 = rev ? -c : c;
        const float p1x = r00 * tl[0] + r01 * tl[1], p1y = r10 * tl[0] + r11 * tl[1];
        const float p2x = r00 * tl[2] + r01 * tl[3], p2y = r10 * tl[2] + r11 * tl[3];
        const float mx = (p2x + p1x) / 2, my = (p2y + p1y) / 2;
        out[(size_t)rev] = Mat23{r00, r01, cx - mx, r10, r11, cy - my};
    }
    return out;
}
inline LineArray transform(const LineArray& lines, const Mat23& T) {    // math.h:341-344
    LineArray out(lines.size());
    for (size_t i = 0; i + 1 < lines.size(); i += 2) {
        const float x = lines[i], y = lines[i + 1];
        out[i] = (T[0] * x + T[1] * y) + T[2];
        out[i + 1] = (T[3] * x + T[4] * y) + T[5];
    }
    return out;
}
inline Mat23 combine(const Point2& translation, const Mat23& T) {       // math.h:427-432
    return Mat23{T[0], T[1], T[2] + translation.x, T[3], T[4], T[5] + translation.y};
}
}   // namespace core

// ConcentricRangeStrategy (searchstrategies/concentricrange.h:35-60)
struct ConcentricRangeStrategy {
    size_t max_tmpl_lines, max_scene_lines;
    Point2 center_position;
    float low_boundary, high_boundary;
    size_t getMaxTmplLines() const noexcept { return max_tmpl_lines; }
    size_t getMaxSceneLines() const noexcept { return max_scene_lines; }
    Point2 getCenterPosition() const noexcept { return center_position; }
    float getLowBoundary() const noexcept { return low_boundary; }
    float getHighBoundary() const noexcept { return high_boundary; }
};
inline std::vector<SearchCombination> establishSearchStrategy(const ConcentricRangeStrategy& s, const LineArray& tmpl, const LineArray& scene) {
    const int32_t cap = (int32_t)(s.max_tmpl_lines * s.max_scene_lines) + 1;
    std::vector<int32_t> pairs((size_t)cap * 2);
    int32_t n = 0;
    check(fdcm_concentric_search(tmpl.data(), (int32_t)(tmpl.size() / 4), scene.data(), (int32_t)(scene.size() / 4), (int32_t)s.max_tmpl_lines,
                                 (int32_t)s.max_scene_lines, s.center_position.x, s.center_position.y, s.low_boundary, s.high_boundary,
                                 pairs.data(), cap, &n));
    std::vector<SearchCombination> out((size_t)n);
    for (int32_t i = 0; i < n; ++i) out[(size_t)i] = {(size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]};
    return out;
}

class FeatureMap;
class SearchStrategy;
class OptimizeStrategy;

// The Models call the concept functions UNQUALIFIED from these helpers, so that overloads living in the namespace of a
// user's type are found by argument-dependent lookup (a Model's own member of the same name would hide them).
namespace adl {
template <class T> Size feature_size(const T& o) { return getFeatureSize(o); }
template <class T> std::array<float, 2> minmax(const T& o, const LineArray& t, const Point2& v) { return minmaxTranslation(o, t, v); }
template <class T> std::vector<std::vector<float>> eval(const T& o, const std::vector<LineArray>& t, const std::vector<std::vector<Point2>>& tr) {
    return evaluate(o, t, tr);
}
template <class T> std::vector<SearchCombination> establish(const T& o, const LineArray& t, const LineArray& s) { return establishSearchStrategy(o, t, s); }
template <class T, class F> std::vector<std::optional<OptimalTranslation>> opt(const T& o, const std::vector<LineArray>& t, const std::vector<Point2>& a, const F& fm) {
    return optimize(o, t, a, fm);
}
template <class T, class S, class O, class F> std::vector<Match> srch(const T& o, const S& s, const O& op, const F& fm, const std::vector<LineArray>& t,
                                                                      const LineArray& sc) {
    return search(o, s, op, fm, t, sc);
}
template <class T> std::vector<Match> pen(const T& o, const std::vector<Match>& m, const std::vector<float>& l) { return penalize(o, m, l); }
}   // namespace adl

// ================================================================================================================
// FeatureMap (featuremap.h:57-124)
// ================================================================================================================
namespace detail {
struct FeatureMapConcept {
    virtual ~FeatureMapConcept() noexcept = default;
    virtual std::unique_ptr<FeatureMapConcept> clone() const = 0;
    virtual Size getFeatureSize() const = 0;
    virtual std::array<float, 2> minmaxTranslation(const LineArray& tmpl, const Point2& align_vec) const = 0;
    virtual std::vector<std::vector<float>> evaluate(const std::vector<LineArray>& templates,
                                                     const std::vector<std::vector<Point2>>& translations) const = 0;
    virtual const Dt3Cuda* dt3cuda() const noexcept { return nullptr; }
};
template <class T>
struct FeatureMapModel final : FeatureMapConcept {
    explicit FeatureMapModel(T value) noexcept : object{std::move(value)} {}
    std::unique_ptr<FeatureMapConcept> clone() const override { return std::make_unique<FeatureMapModel<T>>(*this); }
    Size getFeatureSize() const override { return adl::feature_size(object); }
    std::array<float, 2> minmaxTranslation(const LineArray& tmpl, const Point2& align_vec) const override {
        return adl::minmax(object, tmpl, align_vec);
    }
    std::vector<std::vector<float>> evaluate(const std::vector<LineArray>& templates,
                                             const std::vector<std::vector<Point2>>& translations) const override {
        return adl::eval(object, templates, translations);
    }
    const Dt3Cuda* dt3cuda() const noexcept override {
        if constexpr (std::is_same_v<T, Dt3Cuda>) return &object;
        return nullptr;
    }
    T object;
};
}   // namespace detail

class FeatureMap {
    std::unique_ptr<detail::FeatureMapConcept> pimpl;

public:
    template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, FeatureMap>>>
    /* implicit */ FeatureMap(T const& x) : pimpl{std::make_unique<detail::FeatureMapModel<T>>(x)} {}
    FeatureMap(FeatureMap const& other) : pimpl{other.pimpl->clone()} {}
    FeatureMap& operator=(FeatureMap const& other) { pimpl = other.pimpl->clone(); return *this; }
    FeatureMap(FeatureMap&&) noexcept = default;
    FeatureMap& operator=(FeatureMap&&) noexcept = default;
    Size getFeatureSize() const { return pimpl->getFeatureSize(); }
    std::array<float, 2> minmaxTranslation(const LineArray& tmpl, const Point2& align_vec) const { return pimpl->minmaxTranslation(tmpl, align_vec); }
    std::vector<std::vector<float>> evaluate(const std::vector<LineArray>& templates, const std::vector<std::vector<Point2>>& translations) const {
        return pimpl->evaluate(templates, translations);
    }
    const Dt3Cuda* dt3cuda() const noexcept { return pimpl->dt3cuda(); }   // the device map inside, if that is what it holds
};
inline Size getFeatureSize(const FeatureMap& fm) { return fm.getFeatureSize(); }
inline std::array<float, 2> minmaxTranslation(const FeatureMap& fm, const LineArray& tmpl, const Point2& align_vec) {
    return fm.minmaxTranslation(tmpl, align_vec);
}
inline std::vector<std::vector<float>> evaluate(const FeatureMap& fm, const std::vector<LineArray>& templates,
                                                const std::vector<std::vector<Point2>>& translations) {
    return fm.evaluate(templates, translations);
}

// ================================================================================================================
// SearchStrategy (searchstrategy.h:72-124)
// ================================================================================================================
namespace detail {
struct SearchConcept {
    virtual ~SearchConcept() noexcept = default;
    virtual std::unique_ptr<SearchConcept> clone() const = 0;
    virtual std::vector<SearchCombination> establishSearchStrategy(const LineArray& tmpl, const LineArray& scene) const = 0;
    virtual const DefaultSearch* default_search() const noexcept { return nullptr; }
    virtual const ConcentricRangeStrategy* concentric() const noexcept { return nullptr; }
};
template <class T>
struct SearchModel final : SearchConcept {
    explicit SearchModel(T value) noexcept : object{std::move(value)} {}
    std::unique_ptr<SearchConcept> clone() const override { return std::make_unique<SearchModel<T>>(*this); }
    std::vector<SearchCombination> establishSearchStrategy(const LineArray& tmpl, const LineArray& scene) const override {
        return adl::establish(object, tmpl, scene);
    }
    const DefaultSearch* default_search() const noexcept override {
        if constexpr (std::is_same_v<T, DefaultSearch>) return &object;
        return nullptr;
    }
    const ConcentricRangeStrategy* concentric() const noexcept override {
        if constexpr (std::is_same_v<T, ConcentricRangeStrategy>) return &object;
        return nullptr;
    }
    T object;
};
}   // namespace detail

class SearchStrategy {
    std::unique_ptr<detail::SearchConcept> pimpl;

public:
    template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, SearchStrategy>>>
    /* implicit */ SearchStrategy(T const& x) : pimpl{std::make_unique<detail::SearchModel<T>>(x)} {}
    SearchStrategy(SearchStrategy const& other) : pimpl{other.pimpl->clone()} {}
    SearchStrategy& operator=(SearchStrategy const& other) { pimpl = other.pimpl->clone(); return *this; }
    SearchStrategy(SearchStrategy&&) noexcept = default;
    SearchStrategy& operator=(SearchStrategy&&) noexcept = default;
    std::vector<SearchCombination> establishSearchStrategy(const LineArray& tmpl, const LineArray& scene) const {
        return pimpl->establishSearchStrategy(tmpl, scene);
    }
    const DefaultSearch* default_search() const noexcept { return pimpl->default_search(); }
    const ConcentricRangeStrategy* concentric() const noexcept { return pimpl->concentric(); }
};
inline std::vector<SearchCombination> establishSearchStrategy(const SearchStrategy& s, const LineArray& tmpl, const LineArray& scene) {
    return s.establishSearchStrategy(tmpl, scene);
}

// ================================================================================================================
// OptimizeStrategy (optimizestrategy.h:66-118): optimize(optimizer, templates, alignments, FeatureMap)
// ================================================================================================================
// the CUDA optimisers against an ERASED feature map: they need the device map inside it
inline std::vector<std::optional<OptimalTranslation>> optimize(const BatchOptimize& o, const std::vector<LineArray>& templates,
                                                               const std::vector<Point2>& alignments, const FeatureMap& fm) {
    if (const Dt3Cuda* d = fm.dt3cuda()) return optimize(o, templates, alignments, *d);
    throw std::invalid_argument("BatchOptimize (CUDA) needs a Dt3Cuda feature map");
}
inline std::vector<std::optional<OptimalTranslation>> optimize(const DefaultOptimize& o, const std::vector<LineArray>& templates,
                                                               const std::vector<Point2>& alignments, const FeatureMap& fm) {
    if (const Dt3Cuda* d = fm.dt3cuda()) return optimize(o, templates, alignments, *d);
    throw std::invalid_argument("DefaultOptimize (CUDA) needs a Dt3Cuda feature map");
}

namespace detail {
struct OptimizerConcept {
    virtual ~OptimizerConcept() noexcept = default;
    virtual std::unique_ptr<OptimizerConcept> clone() const = 0;
    virtual std::vector<std::optional<OptimalTranslation>> optimize(const std::vector<LineArray>& templates, const std::vector<Point2>& alignments,
                                                                    const FeatureMap& featuremap) const = 0;
    virtual int batch_size() const noexcept { return -1; }   // >= 0: one of the CUDA optimisers (0 = DefaultOptimize)
};
template <class T>
struct OptimizerModel final : OptimizerConcept {
    explicit OptimizerModel(T value) noexcept : object{std::move(value)} {}
    std::unique_ptr<OptimizerConcept> clone() const override { return std::make_unique<OptimizerModel<T>>(*this); }
    std::vector<std::optional<OptimalTranslation>> optimize(const std::vector<LineArray>& templates, const std::vector<Point2>& alignments,
                                                            const FeatureMap& featuremap) const override {
        return adl::opt(object, templates, alignments, featuremap);
    }
    int batch_size() const noexcept override {
        if constexpr (std::is_same_v<T, BatchOptimize>) return (int)object.batchSize;
        if constexpr (std::is_same_v<T, DefaultOptimize>) return 0;
        return -1;
    }
    T object;
};
}   // namespace detail

class OptimizeStrategy {
    std::unique_ptr<detail::OptimizerConcept> pimpl;

public:
    template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, OptimizeStrategy>>>
    /* implicit */ OptimizeStrategy(T const& x) : pimpl{std::make_unique<detail::OptimizerModel<T>>(x)} {}
    OptimizeStrategy(OptimizeStrategy const& other) : pimpl{other.pimpl->clone()} {}
    OptimizeStrategy& operator=(OptimizeStrategy const& other) { pimpl = other.pimpl->clone(); return *this; }
    OptimizeStrategy(OptimizeStrategy&&) noexcept = default;
    OptimizeStrategy& operator=(OptimizeStrategy&&) noexcept = default;
    std::vector<std::optional<OptimalTranslation>> optimize(const std::vector<LineArray>& templates, const std::vector<Point2>& alignments,
                                                            const FeatureMap& featuremap) const {
        return pimpl->optimize(templates, alignments, featuremap);
    }
    int batch_size() const noexcept { return pimpl->batch_size(); }
};
inline std::vector<std::optional<OptimalTranslation>> optimize(const OptimizeStrategy& o, const std::vector<LineArray>& templates,
                                                               const std::vector<Point2>& alignments, const FeatureMap& fm) {
    return o.optimize(templates, alignments, fm);
}

// ================================================================================================================
// MatchStrategy (matchstrategy.h:83-140) and the DefaultMatch driver (defaultmatch.cpp:32-89)
// ================================================================================================================
inline std::vector<Match> search(const DefaultMatch&, const SearchStrategy& searcher, const OptimizeStrategy& optimizer, const FeatureMap& featuremap,
                                 const std::vector<LineArray>& templates, const LineArray& originalScene) {
    const Size fs = featuremap.getFeatureSize();
    if (templates.empty() || originalScene.size() < 4 || (fs.x == 0 && fs.y == 0)) return {};   // defaultmatch.cpp:40-41

    // ---- fused device path: everything the GPU kernels cover, in one launch ----
    const Dt3Cuda* dmap = featuremap.dt3cuda();
    const int batch = optimizer.batch_size();
    if (dmap && batch >= 0 && (searcher.default_search() || searcher.concentric())) {
        if (const DefaultSearch* ds = searcher.default_search())
            return detail::search(batch, *ds, *dmap, templates, originalScene, FDCM_PENALTY_NONE, 0.f, 0);
        const ConcentricRangeStrategy* cs = searcher.concentric();
        return detail::search_concentric(batch, *cs, *dmap, templates, originalScene);
    }

    // ---- generic composition: the reference's loop over the erased interfaces ----
    std::vector<LineArray> aligned_templates;
    std::vector<int> template_indices;
    std::vector<Point2> alignments;
    std::vector<Mat23> transforms;
    for (size_t tmpl_idx = 0; tmpl_idx < templates.size(); ++tmpl_idx) {
        const LineArray& tmpl = templates[tmpl_idx];
        if (tmpl.size() < 4) continue;                                                           // :54
        for (const SearchCombination& c : establishSearchStrategy(searcher, tmpl, originalScene)) {
            const float* scene_line = originalScene.data() + 4 * c.sceneLineIdx;
            const float* tmpl_line = tmpl.data() + 4 * c.tmplLineIdx;
            const Point2 align_vec = core::normalize(scene_line);
            const std::array<Mat23, 2> tr = core::align(tmpl_line, scene_line);
            for (int rev = 0; rev < 2; ++rev) {
                transforms.push_back(tr[(size_t)rev]);
                template_indices.push_back((int)tmpl_idx);
                aligned_templates.push_back(core::transform(tmpl, tr[(size_t)rev]));
                alignments.push_back(align_vec);
            }
        }
    }
    const std::vector<std::optional<OptimalTranslation>> results = optimize(optimizer, aligned_templates, alignments, featuremap);
    std::vector<Match> all_matches;
    for (size_t i = 0; i < aligned_templates.size(); ++i)
        if (results[i].has_value())
            all_matches.push_back(Match{template_indices[i], results[i]->score, core::combine(results[i]->translation, transforms[i])});
    return all_matches;
}

namespace detail {
struct MatcherConcept {
    virtual ~MatcherConcept() noexcept = default;
    virtual std::unique_ptr<MatcherConcept> clone() const = 0;
    virtual std::vector<Match> search(const SearchStrategy& searcher, const OptimizeStrategy& optimizer, const FeatureMap& featuremap,
                                      const std::vector<LineArray>& templates, const LineArray& scene) const = 0;
};
template <class T>
struct MatcherModel final : MatcherConcept {
    explicit MatcherModel(T value) noexcept : object{std::move(value)} {}
    std::unique_ptr<MatcherConcept> clone() const override { return std::make_unique<MatcherModel<T>>(*this); }
    std::vector<Match> search(const SearchStrategy& searcher, const OptimizeStrategy& optimizer, const FeatureMap& featuremap,
                              const std::vector<LineArray>& templates, const LineArray& scene) const override {
        return adl::srch(object, searcher, optimizer, featuremap, templates, scene);
    }
    T object;
};
}   // namespace detail

class MatchStrategy {
    std::unique_ptr<detail::MatcherConcept> pimpl;

public:
    template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, MatchStrategy>>>
    /* implicit */ MatchStrategy(T const& x) : pimpl{std::make_unique<detail::MatcherModel<T>>(x)} {}
    MatchStrategy(MatchStrategy const& other) : pimpl{other.pimpl->clone()} {}
    MatchStrategy& operator=(MatchStrategy const& other) { pimpl = other.pimpl->clone(); return *this; }
    MatchStrategy(MatchStrategy&&) noexcept = default;
    MatchStrategy& operator=(MatchStrategy&&) noexcept = default;
    std::vector<Match> search(const SearchStrategy& searcher, const OptimizeStrategy& optimizer, const FeatureMap& featuremap,
                              const std::vector<LineArray>& templates, const LineArray& scene) const {
        return pimpl->search(searcher, optimizer, featuremap, templates, scene);
    }
};
inline std::vector<Match> search(const MatchStrategy& matcher, const SearchStrategy& searcher, const OptimizeStrategy& optimizer,
                                 const FeatureMap& featuremap, const std::vector<LineArray>& templates, const LineArray& scene) {
    return matcher.search(searcher, optimizer, featuremap, templates, scene);
}

// ================================================================================================================
// PenaltyStrategy (penaltystrategy.h:60-107)
// ================================================================================================================
namespace detail {
struct PenaltyConcept {
    virtual ~PenaltyConcept() noexcept = default;
    virtual std::unique_ptr<PenaltyConcept> clone() const = 0;
    virtual std::vector<Match> penalize(const std::vector<Match>& matches, const std::vector<float>& templatelengths) const = 0;
};
template <class T>
struct PenaltyModel final : PenaltyConcept {
    explicit PenaltyModel(T value) noexcept : object{std::move(value)} {}
    std::unique_ptr<PenaltyConcept> clone() const override { return std::make_unique<PenaltyModel<T>>(*this); }
    std::vector<Match> penalize(const std::vector<Match>& matches, const std::vector<float>& templatelengths) const override {
        return adl::pen(object, matches, templatelengths);
    }
    T object;
};
}   // namespace detail

class PenaltyStrategy {
    std::unique_ptr<detail::PenaltyConcept> pimpl;

public:
    template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, PenaltyStrategy>>>
    /* implicit */ PenaltyStrategy(T const& x) : pimpl{std::make_unique<detail::PenaltyModel<T>>(x)} {}
    PenaltyStrategy(PenaltyStrategy const& other) : pimpl{other.pimpl->clone()} {}
    PenaltyStrategy& operator=(PenaltyStrategy const& other) { pimpl = other.pimpl->clone(); return *this; }
    PenaltyStrategy(PenaltyStrategy&&) noexcept = default;
    PenaltyStrategy& operator=(PenaltyStrategy&&) noexcept = default;
    std::vector<Match> penalize(const std::vector<Match>& matches, const std::vector<float>& templatelengths) const {
        return pimpl->penalize(matches, templatelengths);
    }
};
inline std::vector<Match> penalize(const PenaltyStrategy& p, const std::vector<Match>& matches, const std::vector<float>& templatelengths) {
    return p.penalize(matches, templatelengths);
}

}   // namespace openfdcm::cuda
