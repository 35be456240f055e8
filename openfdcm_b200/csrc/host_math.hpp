// host_math.hpp — host-side scalar geometry of the hot paths (scene shift, angle keys, orientation
// bins, slope-threshold table, propagation schedule, DefaultSearch ordering, bounds).
// These are the O(#lines) pieces the C ABI runs on the host (with the host libm, exactly like the
// reference does) before the device takes over.  `path:line` citations are relative to the
// reference tree.  Compiled without FMA contraction.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace fdcm {

constexpr float kPIf = 3.14159265358979323846f;     // core/math.h:38-40
constexpr float kPI_2f = 1.57079632679489661923f;   // core/math.h:44-46
constexpr int kMaxDepth = 64;

// ---- float <-> totally ordered int (for bisection over the float line) -----------------------
inline int32_t float_to_ordered(float f) {
    int32_t i;
    std::memcpy(&i, &f, 4);
    return i >= 0 ? i : (int32_t)(0x80000000u - (uint32_t)i);   // -0.0 -> 0 as well; fine for bisection
}
inline float ordered_to_float(int32_t o) {
    int32_t i = o >= 0 ? o : (int32_t)(0x80000000u - (uint32_t)o);
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}

// angle keys (dt3cpu.h:188-190), ascending = plane order (std::set order)
inline std::vector<float> angle_keys(int depth) {
    std::vector<float> k;
    k.reserve(depth);
    for (int i = 0; i < depth; ++i) k.push_back(float(i) * kPIf / float(depth) - kPI_2f);
    std::sort(k.begin(), k.end());
    k.erase(std::unique(k.begin(), k.end()), k.end());
    return k;
}

// closestOrientation on an angle (dt3cpu.h:93-114)
inline int closest_orientation_angle(const float* keys, int n, float a) {
    const int it = (int)(std::upper_bound(keys, keys + n, a) - keys);
    if (it != n && it != 0) {
        const float ud = std::abs(a - keys[it]);
        const float ld = std::abs(a - keys[it - 1]);
        return ld < ud ? it - 1 : it;
    }
    const float a1 = a - keys[0];
    const float a2 = a - keys[n - 1];
    if (std::min(a1, std::abs(a1 - kPIf)) < std::min(a2, std::abs(a2 - kPIf))) return 0;
    return n - 1;
}

// bin of a slope dy/dx: getAngle (math.h:295-299) + closestOrientation, with the host libm atanf
inline int bin_of_slope(const float* keys, int n, float slope) {
    return closest_orientation_angle(keys, n, std::atan(slope));
}
inline int bin_of_line(const float* keys, int n, const float* l) {
    return bin_of_slope(keys, n, (l[3] - l[1]) / (l[2] - l[0]));
}

// Slope-threshold table: bin(slope) is piecewise constant along the ordered float line (atanf is
// monotone); the device reproduces the host's atanf-based choice with one IEEE division and a
// binary search.  thr[j] = smallest slope (as float) of piece j+1; piece_bin[j] = bin of piece j.
struct SlopeTable {
    std::vector<float> thr;       // ascending, size = pieces-1
    std::vector<int32_t> piece_bin;
    int32_t nan_bin = 0;
};

inline void slope_table_split(const float* keys, int n, int32_t lo, int blo, int32_t hi, int bhi,
                              std::vector<std::pair<int32_t, int>>& changes) {
    // invariant: bin(lo) = blo != bhi = bin(hi), lo < hi (ordered ints)
    if (hi - lo == 1) {
        changes.emplace_back(hi, bhi);
        return;
    }
    const int32_t mid = lo + (int32_t)(((int64_t)hi - (int64_t)lo) / 2);
    const int bm = bin_of_slope(keys, n, ordered_to_float(mid));
    if (bm != blo) slope_table_split(keys, n, lo, blo, mid, bm, changes);
    if (bm != bhi) slope_table_split(keys, n, mid, bm, hi, bhi, changes);
}

inline SlopeTable build_slope_table(const float* keys, int n) {
    SlopeTable t;
    const float inf = std::numeric_limits<float>::infinity();
    t.nan_bin = bin_of_slope(keys, n, std::numeric_limits<float>::quiet_NaN());
    // anchors: -inf, a grid of slopes, +inf (the wrap-around makes bin(-inf) == bin(+inf) possible,
    // so interior anchors are required; tan of the key mid-points guarantees one anchor per piece)
    std::vector<float> anchors{-inf, 0.f, inf};
    for (int i = 0; i < n; ++i) {
        anchors.push_back(std::tan((double)keys[i]) > 1e30 ? 1e30f : (float)std::tan((double)keys[i]));
        if (i + 1 < n) anchors.push_back((float)std::tan(0.5 * ((double)keys[i] + (double)keys[i + 1])));
    }
    for (float& a : anchors)
        if (std::isnan(a)) a = 0.f;
    std::sort(anchors.begin(), anchors.end());
    std::vector<std::pair<int32_t, int>> changes;   // (ordered slope where a new piece starts, its bin)
    int32_t prev = float_to_ordered(anchors[0]);
    int bprev = bin_of_slope(keys, n, anchors[0]);
    t.piece_bin.push_back(bprev);
    for (size_t i = 1; i < anchors.size(); ++i) {
        const int32_t cur = float_to_ordered(anchors[i]);
        if (cur == prev) continue;
        const int bcur = bin_of_slope(keys, n, anchors[i]);
        if (bcur != bprev) slope_table_split(keys, n, prev, bprev, cur, bcur, changes);
        prev = cur;
        bprev = bcur;
    }
    std::sort(changes.begin(), changes.end());
    for (auto& c : changes) {
        t.thr.push_back(ordered_to_float(c.first));
        t.piece_bin.push_back(c.second);
    }
    return t;
}

inline int slope_table_lookup(const SlopeTable& t, float slope) {
    if (std::isnan(slope)) return t.nan_bin;
    // number of thresholds <= slope
    const int j = (int)(std::upper_bound(t.thr.begin(), t.thr.end(), slope) - t.thr.begin());
    return t.piece_bin[j];
}

// getSceneCenteredTranslation (dt3cpu.cpp:109-116)
inline void scene_centered_translation(const float* scene, int64_t n, float padding, float shift[2], int64_t size[2]) {
    float mn[2] = {scene[0], scene[1]}, mx[2] = {scene[0], scene[1]};
    for (int64_t i = 0; i < 2 * n; ++i)
        for (int a = 0; a < 2; ++a) {
            const float v = scene[2 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1];
    const float ratio = std::max(1.f, padding);
    const float req = ratio * std::max(dx, dy) * 1.f;
    for (int a = 0; a < 2; ++a) {
        shift[a] = req / 2.f - (mx[a] + mn[a]) / 2.f;
        size[a] = (int64_t)(size_t)std::ceil(req + 1.f);
    }
}

// rasterizeVector (core/drawing.h:57-67).  The reference evaluates `x - 2.0*cond*x` in double:
// for x != 0 that is exactly +-x, for x == +-0 it is always +0 (and NaN stays NaN).
inline float signed_pick(float x, bool negate) { return x == 0.f ? 0.f : (negate ? -x : x); }
inline void rasterize_vector(float vx, float vy, float& ox, float& oy) {
    const float t = vy / vx;
    if (t >= -1.0f && t < 1.0f) {
        const bool c = vx < 0;
        ox = c ? -1.f : 1.f;
        oy = signed_pick(t, c);
        return;
    }
    const bool c = vy < 0;
    const float inv = 1.f / t;
    ox = signed_pick(inv, c);
    oy = c ? -1.f : 1.f;
}

// propagateOrientation schedule (dt3cpu.cpp:77-107): the 4*m (c1, c2, weight) steps
struct PropStep { int32_t c1, c2; float w; };
inline std::vector<PropStep> propagation_schedule(const std::vector<float>& keys, float coeff) {
    std::vector<PropStep> s;
    const int m = (int)keys.size();
    if (m == 0) return s;
    const int fwd = static_cast<int>(std::ceil(1.5 * m));
    const int bwd = -static_cast<int>(std::floor(1.5 * m));
    auto run = [&](int start, int end, int step) {
        for (int c = start; c != end; c += step) {
            const int c1 = (m + ((c - step) % m)) % m;
            const int c2 = (m + (c % m)) % m;
            const float h = std::abs(keys[c1] - keys[c2]);
            const float min_h = std::min(h, std::abs(h - kPIf));
            s.push_back(PropStep{c1, c2, coeff * min_h});
        }
    };
    run(0, fwd, 1);
    run(m, bwd, -1);
    return s;
}

// lineIntegral direction of a plane (core/imgproc.h:41-48): rasterised direction + sweep mode
struct IntegralDir { float rx, ry; int32_t mode; };   // mode 1: x-major (|rx|==1), 2: y-major, 0: none
inline IntegralDir integral_direction(float angle) {
    IntegralDir d;
    rasterize_vector(std::cos(angle), std::sin(angle), d.rx, d.ry);
    d.mode = (std::abs(d.rx) == 1) ? 1 : ((std::abs(d.ry) == 1) ? 2 : 0);
    return d;
}

// Eigen 3.4.0 float redux order (see oracle / SURVEY App. A.14), used for getTemplateLengths (math.h:319-324)
inline float eigen_sum(const float* c, int64_t n) {
    if (n == 0) return 0.f;
    const int64_t n4 = (n / 4) * 4, n8 = (n / 8) * 8;
    if (n4 == 0) {
        float r = c[0];
        for (int64_t i = 1; i < n; ++i) r = r + c[i];
        return r;
    }
    float A0 = c[0], A1 = c[1], A2 = c[2], A3 = c[3];
    if (n4 > 4) {
        float B0 = c[4], B1 = c[5], B2 = c[6], B3 = c[7];
        for (int64_t i = 8; i < n8; i += 8) {
            A0 = A0 + c[i]; A1 = A1 + c[i + 1]; A2 = A2 + c[i + 2]; A3 = A3 + c[i + 3];
            B0 = B0 + c[i + 4]; B1 = B1 + c[i + 5]; B2 = B2 + c[i + 6]; B3 = B3 + c[i + 7];
        }
        A0 = A0 + B0; A1 = A1 + B1; A2 = A2 + B2; A3 = A3 + B3;
        if (n4 > n8) { A0 = A0 + c[n8]; A1 = A1 + c[n8 + 1]; A2 = A2 + c[n8 + 2]; A3 = A3 + c[n8 + 3]; }
    }
    float r = (A0 + A2) + (A1 + A3);
    for (int64_t i = n4; i < n; ++i) r = r + c[i];
    return r;
}

inline float line_length(const float* l) {   // math.h:306-308
    const float dx = l[2] - l[0], dy = l[3] - l[1];
    return std::sqrt(dx * dx + dy * dy);
}

// argsort descending with the reference's exact comparator and std::sort (math.h:107-116 + std::greater)
inline std::vector<long> argsort_desc(const float* v, int64_t n) {
    std::vector<long> ind((size_t)n);
    for (int64_t i = 0; i < n; ++i) ind[(size_t)i] = (long)i;
    std::sort(ind.begin(), ind.end(), [v](long const i1, long const i2) { return v[i1] > v[i2]; });
    return ind;
}

// binarySearch with std::greater on a descending array (math.h:138-146)
inline int64_t closest_in_descending(const float* s, int64_t n, float value) {
    const float* it = std::lower_bound(s, s + n, value, std::greater<float>());
    if (it == s) return 0;
    if (it == s + n) return (it - 1) - s;
    return std::abs(value - *it) < std::abs(value - *(it - 1)) ? it - s : (it - 1) - s;
}

// getCenteredRange (searchstrategies/defaultsearch.h:40-47)
inline void centered_range(int64_t center, int64_t vec_size, int64_t max_length, int64_t& b, int64_t& e) {
    b = std::max(0, int(center) - int(max_length / 2));
    e = std::min<int64_t>(b + max_length, vec_size);
    b = std::max(0, int(e) - int(max_length));
}

// minmaxTranslation (dt3cpu.cpp:30-75) on the host (FeatureMap concept entry point)
inline void minmax_translation(const float* tmpl, int64_t n, const float vec[2], const int64_t fsize[2],
                               const float extra[2], float out[2]) {
    const float inf = std::numeric_limits<float>::infinity();
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if (std::fabs(vec[0]) <= 1e-5f && std::fabs(vec[1]) <= 1e-5f) { out[0] = out[1] = inf; return; }
    float mn[2] = {0, 0}, mx[2] = {0, 0};
    if (n > 0) { mn[0] = mx[0] = tmpl[0]; mn[1] = mx[1] = tmpl[1]; }
    for (int64_t i = 0; i < 2 * n; ++i)
        for (int a = 0; a < 2; ++a) {
            const float v = tmpl[2 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    float E[2][2];
    for (int a = 0; a < 2; ++a) {
        mn[a] += extra[a];
        mx[a] += extra[a];
    }
    const float size[2] = {(float)fsize[0], (float)fsize[1]};
    if ((size[0] - 1 - mx[0]) < 0 || (size[1] - 1 - mx[1]) < 0 || mn[0] < 0 || mn[1] < 0) { out[0] = out[1] = nan; return; }
    for (int a = 0; a < 2; ++a) {
        const float m[4] = {(-mx[a]) / vec[a], (-mn[a]) / vec[a], (size[a] - mx[a] - 1.f) / vec[a], (size[a] - mn[a] - 1.f) / vec[a]};
        float neg_max = -inf, pos_min = inf;
        bool neg_nan = false, pos_nan = false;
        for (int j = 0; j < 4; ++j) {
            const bool sg = std::signbit(m[j]);
            const float pos = sg ? inf : m[j];
            const float neg = sg ? m[j] : -inf;
            pos_nan |= std::isnan(pos);
            neg_nan |= std::isnan(neg);
            if (neg > neg_max) neg_max = neg;
            if (pos < pos_min) pos_min = pos;
        }
        E[0][a] = neg_nan ? nan : neg_max;
        E[1][a] = pos_nan ? nan : pos_min;
    }
    const bool f0 = std::isfinite(E[0][0]) && std::isfinite(E[1][0]);
    const bool f1 = std::isfinite(E[0][1]) && std::isfinite(E[1][1]);
    if (f0 && f1) { out[0] = std::max(E[0][0], E[0][1]); out[1] = std::min(E[1][0], E[1][1]); }
    else if (f0) { out[0] = E[0][0]; out[1] = E[1][0]; }
    else { out[0] = E[0][1]; out[1] = E[1][1]; }
}

}   // namespace fdcm
