// fdcm_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A from-scratch, Eigen-free restatement of the reference pipeline
//   Dt3Cpu + DefaultSearch + {Batch,Default}Optimize + DefaultMatch + penalize + sort
// of Innoptech/OpenFDCM v0.10.0.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product path
// (openfdcm_b200/csrc) never links, imports or calls it.
//
// Parity status: the reference itself cannot be compiled here (Eigen 3.4.0, BS::thread_pool,
// packio are network FetchContent dependencies, absent on disk), so this oracle is pinned
// against the reference's own known-answer tests (tests/test_oracle_golden.py re-expresses
// every hot-path row of SURVEY.md §4.2).  Behaviour the reference tests do NOT pin (the
// in-place aliasing of the L2 column pass, Eigen's LinSpaced / redux summation order, atanf at
// bin boundaries) follows the reference source text literally: "parity unpinned by reference
// tests" for those items.
//
// Build: g++ -O3 -fno-math-errno -ffp-contract=off (mirrors reference CMakeLists.txt:29-32:
// x86-64 SSE2 baseline, no FMA, asserts on).  No -march=native, no -ffast-math.
//
// All `path:line` citations are relative to the reference tree (modules/...).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <thread>
#include <vector>

namespace orc {

using Index = long;
static const float kPIf   = 3.14159265358979323846f;   // core/math.h:38-40 (M_PIf)
static const float kPI_2f = 1.57079632679489661923f;   // core/math.h:44-46 (M_PI_2f)
static const float kInf   = std::numeric_limits<float>::infinity();

// Column-major image like core::RawImage<float> (core/math.h:57): element (y,x) at d[x*rows+y].
struct Img {
    Index rows{0}, cols{0};
    std::vector<float> d;
    Img() = default;
    Img(Index r, Index c, float fill) : rows(r), cols(c), d((size_t)r * (size_t)c, fill) {}
    inline float& at(Index y, Index x) { return d[(size_t)x * rows + y]; }
    inline float at(Index y, Index x) const { return d[(size_t)x * rows + y]; }
};

static Img transposed(const Img& a) {   // imgproc.h:181 `.transpose().eval()`
    Img t(a.cols, a.rows, 0.f);
    const Index B = 32;
    for (Index x0 = 0; x0 < a.cols; x0 += B)
        for (Index y0 = 0; y0 < a.rows; y0 += B)
            for (Index x = x0; x < std::min(a.cols, x0 + B); ++x)
                for (Index y = y0; y < std::min(a.rows, y0 + B); ++y)
                    t.d[(size_t)y * t.rows + x] = a.d[(size_t)x * a.rows + y];
    return t;
}

// ---------------------------------------------------------------------------------------------
// simple parallel-for (stands in for BS::thread_pool::submit_task + wait; one task per item,
// dt3cpu.h:209-218, batchoptimize.cpp:102-114)
// ---------------------------------------------------------------------------------------------
template <class F>
static void parallel_for(size_t n, int nthreads, F&& f) {
    if (nthreads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    int nt = (int)std::min<size_t>(n, (size_t)nthreads);
    th.reserve(nt);
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&] {
            for (;;) {
                size_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto& t : th) t.join();
}

// ---------------------------------------------------------------------------------------------
// core/math.h helpers
// ---------------------------------------------------------------------------------------------
// relativelyEqual(a, b) with defaults rtol=1e-10, atol=epsilon<float> (math.h:183-189); the
// comparison is carried out in double exactly as the reference's mixed float/double expression.
static inline bool relativelyEqual(float a, float b, double rtol = 1e-10,
                                   double atol = (double)std::numeric_limits<float>::epsilon()) {
    return std::fabs(a - b) <= atol + rtol * std::max(std::fabs(a), std::fabs(b));
}

// Eigen 3.4.0 dynamic-size float `.sum()` (Redux.h LinearVectorizedTraversal, SSE2 Packet4f,
// 2x unrolled, predux = (a0+a2)+(a1+a3)).  Used by dt3cpu.cpp:175 and math.h:322.
static float eigen_sum(const float* c, Index n) {
    if (n == 0) return 0.f;
    const Index n4 = (n / 4) * 4, n8 = (n / 8) * 8;
    if (n4 == 0) {
        float r = c[0];
        for (Index i = 1; i < n; ++i) r = r + c[i];
        return r;
    }
    float A[4] = {c[0], c[1], c[2], c[3]};
    if (n4 > 4) {
        float B[4] = {c[4], c[5], c[6], c[7]};
        for (Index i = 8; i < n8; i += 8)
            for (int j = 0; j < 4; ++j) {
                A[j] = A[j] + c[i + j];
                B[j] = B[j] + c[i + 4 + j];
            }
        for (int j = 0; j < 4; ++j) A[j] = A[j] + B[j];
        if (n4 > n8)
            for (int j = 0; j < 4; ++j) A[j] = A[j] + c[n8 + j];
    }
    float r = (A[0] + A[2]) + (A[1] + A[3]);
    for (Index i = n4; i < n; ++i) r = r + c[i];
    return r;
}

static inline float line_length(const float* l) {   // math.h:306-308 colwise().norm()
    const float dx = l[2] - l[0], dy = l[3] - l[1];
    return std::sqrt(dx * dx + dy * dy);
}

// math.h:331-333: colwise().normalized() — Eigen: z = squaredNorm; z>0 ? v/sqrt(z) : v
static inline void unit_vec(const float* l, float& ux, float& uy) {
    const float dx = l[2] - l[0], dy = l[3] - l[1];
    const float z = dx * dx + dy * dy;
    if (z > 0.f) {
        const float n = std::sqrt(z);
        ux = dx / n;
        uy = dy / n;
    } else {
        ux = dx;
        uy = dy;
    }
}

struct Mat23 { float r00, r01, tx, r10, r11, ty; };

// math.h:341-344: (R * p) + t, coefficient-based product r_i0*x + r_i1*y then + t (no FMA).
static inline void transform_pt(const Mat23& T, float x, float y, float& ox, float& oy) {
    ox = (T.r00 * x + T.r01 * y) + T.tx;
    oy = (T.r10 * x + T.r11 * y) + T.ty;
}

// math.h:387-406 align(): the two rigid transforms bringing alignment_line onto ref_line.
static void align(const float* tl, const float* sl, Mat23& T1, Mat23& T2) {
    float tx, ty, ax, ay;
    unit_vec(tl, tx, ty);
    unit_vec(sl, ax, ay);
    const float c = ax * tx + ay * ty;
    const float s = ay * tx - ax * ty;
    const float cx = (sl[2] + sl[0]) / 2, cy = (sl[3] + sl[1]) / 2;   // getCenter(ref_line)
    {
        // rotate(alignment_line, rot1) then getCenter
        const float p1x = c * tl[0] + (-s) * tl[1], p1y = s * tl[0] + c * tl[1];
        const float p2x = c * tl[2] + (-s) * tl[3], p2y = s * tl[2] + c * tl[3];
        const float mx = (p2x + p1x) / 2, my = (p2y + p1y) / 2;
        T1 = Mat23{c, -s, cx - mx, s, c, cy - my};
    }
    {
        const float nc = -c, ns = -s;
        const float p1x = nc * tl[0] + s * tl[1], p1y = ns * tl[0] + nc * tl[1];
        const float p2x = nc * tl[2] + s * tl[3], p2y = ns * tl[2] + nc * tl[3];
        const float mx = (p2x + p1x) / 2, my = (p2y + p1y) / 2;
        T2 = Mat23{nc, s, cx - mx, ns, nc, cy - my};
    }
}

// ---------------------------------------------------------------------------------------------
// core/drawing.{h,cpp}
// ---------------------------------------------------------------------------------------------
static inline void rasterizeVector(float vx, float vy, float& ox, float& oy) {   // drawing.h:57-67
    const float t = vy / vx;
    if (t >= -1.0f && t < 1.0f) {
        const bool c1 = vx < 0;
        ox = (float)(1 - 2 * (int)c1);
        oy = (float)((double)t - 2.0 * (double)c1 * (double)t);
        return;
    }
    const bool c2 = vy < 0;
    const float inv = 1 / t;
    ox = (float)((double)inv - 2.0 * (double)c2 * (double)inv);
    oy = (float)(1 - 2 * (int)c2);
}

// Eigen 3.4.0 LinSpaced<float>(n, lo, hi), scalar functor (NullaryFunctors.h linspaced_op_impl)
struct LinSpaced {
    float lo, hi, step;
    Index size1;
    bool flip;
    LinSpaced(Index n, float lo_, float hi_)
        : lo(lo_), hi(hi_), step(n == 1 ? 0.f : (hi_ - lo_) / (float)(n - 1)), size1(n == 1 ? 1 : n - 1),
          flip(std::fabs(hi_) < std::fabs(lo_)) {}
    inline float operator()(Index i) const {
        if (flip) return (i == 0) ? lo : (hi - (float)(size1 - i) * step);
        return (i == size1) ? hi : (lo + (float)i * step);
    }
};

static inline int outcode(float x, float y, float xmin, float xmax, float ymin, float ymax) {
    int code = 0;   // drawing.cpp:37-50: LEFT=1 RIGHT=2 BOTTOM=4 TOP=8
    if (x < xmin) code |= 1; else if (x > xmax) code |= 2;
    if (y < ymin) code |= 4; else if (y > ymax) code |= 8;
    return code;
}

// drawing.cpp:64-112 Cohen-Sutherland; returns false if the line is purged
static bool clip_line(float* l, float xmin, float xmax, float ymin, float ymax) {
    float &x1 = l[0], &y1 = l[1], &x2 = l[2], &y2 = l[3];
    int c1 = outcode(x1, y1, xmin, xmax, ymin, ymax);
    int c2 = outcode(x2, y2, xmin, xmax, ymin, ymax);
    for (;;) {
        if (c1 == 0 && c2 == 0) return true;
        if (c1 & c2) return false;
        if (c1 != 0) {
            if (c1 & 8)      { x1 = x1 + (x2 - x1) * (ymax - y1) / (y2 - y1); y1 = ymax; }
            else if (c1 & 4) { x1 = x1 + (x2 - x1) * (ymin - y1) / (y2 - y1); y1 = ymin; }
            else if (c1 & 2) { y1 = y1 + (y2 - y1) * (xmax - x1) / (x2 - x1); x1 = xmax; }
            else if (c1 & 1) { y1 = y1 + (y2 - y1) * (xmin - x1) / (x2 - x1); x1 = xmin; }
            c1 = outcode(x1, y1, xmin, xmax, ymin, ymax);
            continue;
        }
        if (c2 & 8)      { x2 = x2 + (x1 - x2) * (ymax - y2) / (y1 - y2); y2 = ymax; }
        else if (c2 & 4) { x2 = x2 + (x1 - x2) * (ymin - y2) / (y1 - y2); y2 = ymin; }
        else if (c2 & 2) { y2 = y2 + (y1 - y2) * (xmax - x2) / (x1 - x2); x2 = xmax; }
        else if (c2 & 1) { y2 = y2 + (y1 - y2) * (xmin - x2) / (x1 - x2); x2 = xmin; }
        c2 = outcode(x2, y2, xmin, xmax, ymin, ymax);
    }
}

// drawing.h:74-102 rasterizeLine → list of (x,y) pixels
static void rasterizeLine(const float* l, std::vector<Index>& xs, std::vector<Index>& ys) {
    xs.clear();
    ys.clear();
    const float p1x = l[0], p1y = l[1], p2x = l[2], p2y = l[3];
    // allClose(p2, p1): |p2-p1| <= 1e-5f + 0.f*|p1| for both components (math.h:203-209)
    if (std::fabs(p2x - p1x) <= (1e-5f + 0.0f * std::fabs(p1x)) &&
        std::fabs(p2y - p1y) <= (1e-5f + 0.0f * std::fabs(p1y))) {
        xs.push_back((Index)std::round(p1x));
        ys.push_back((Index)std::round(p1y));
        return;
    }
    const float lvx = p2x - p1x, lvy = p2y - p1y;
    float rx, ry;
    rasterizeVector(lvx, lvy, rx, ry);
    if (relativelyEqual(rx, 0.0f)) {
        const int size = int(lvy / ry) + 1;
        LinSpaced ly(size, p1y, p2y);
        for (int i = 0; i < size; ++i) {
            xs.push_back((Index)std::round(p1x));
            ys.push_back((Index)std::round(ly(i)));
        }
        return;
    }
    if (relativelyEqual(ry, 0.0f)) {
        const int size = int(lvx / rx) + 1;
        LinSpaced lx(size, p1x, p2x);
        for (int i = 0; i < size; ++i) {
            xs.push_back((Index)std::round(lx(i)));
            ys.push_back((Index)std::round(p1y));
        }
        return;
    }
    const int size = static_cast<int>(std::max(lvx / rx, lvy / ry)) + 1;
    LinSpaced lx(size, p1x, p2x), ly(size, p1y, p2y);
    for (int i = 0; i < size; ++i) {
        xs.push_back((Index)std::round(lx(i)));
        ys.push_back((Index)std::round(ly(i)));
    }
}

// drawing.h:111-125 drawLines (clip to [0,W-1]x[0,H-1], rasterise, set colour)
static void drawLines(Img& img, const float* lines, Index n, float color) {
    if (n == 0) return;
    std::vector<Index> xs, ys;
    for (Index i = 0; i < n; ++i) {
        float l[4] = {lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]};
        if (!clip_line(l, 0.f, static_cast<float>(img.cols - 1), 0.f, static_cast<float>(img.rows - 1))) continue;
        rasterizeLine(l, xs, ys);
        for (size_t k = 0; k < xs.size(); ++k) img.at(ys[k], xs[k]) = color;
    }
}

// ---------------------------------------------------------------------------------------------
// core/imgproc.h
// ---------------------------------------------------------------------------------------------
// imgproc.h:91-130 — literal, including the in-place second loop (reads img(v_k,i) after it may
// already have been overwritten when v_k < q).
static void colPassL2(Img& img) {
    const Index R = img.rows;
    std::vector<Index> sq(R);
    for (Index i = 0; i < R; ++i) sq[i] = i * i;
    std::vector<Index> v(R);
    std::vector<float> z(R + 1);
    for (Index i = 0; i < img.cols; ++i) {
        float* c = &img.d[(size_t)i * R];
        Index k = 0;
        v[0] = 0;
        z[0] = -kInf;
        z[1] = kInf;
        for (Index q = 1; q < R; ++q) {
            while (true) {
                const Index v_k = v[k];
                const float s = (c[q] + sq[q] - c[v_k] - sq[v_k]) / (2 * q - 2 * v_k);
                if (s > z[k]) {
                    ++k;
                    v[k] = q;
                    z[k] = s;
                    z[k + 1] = kInf;
                    break;
                }
                --k;
            }
        }
        k = 0;
        for (Index q = 0; q < R; ++q) {
            while (z[k + 1] < (float)q) ++k;
            const Index v_k = v[k];
            const Index q_ = q - v_k;
            c[q] = c[v_k] + sq[std::labs(q_)];
        }
    }
}

// imgproc.h:137-146
static void colPassL1(Img& img) {
    const Index R = img.rows;
    for (Index q = 1; q < img.cols; ++q) {
        float* a = &img.d[(size_t)q * R];
        const float* b = &img.d[(size_t)(q - 1) * R];
        for (Index y = 0; y < R; ++y) a[y] = std::min(a[y], b[y] + 1);
    }
    for (Index q = img.cols - 2; q >= 0; --q) {
        float* a = &img.d[(size_t)q * R];
        const float* b = &img.d[(size_t)(q + 1) * R];
        for (Index y = 0; y < R; ++y) a[y] = std::min(a[y], b[y] + 1);
    }
}

enum Distance { L2 = 0, L2_SQUARED = 1, L1 = 2 };   // imgproc.h:148

// imgproc.h:169-194; size = (W, H)
static Img distanceTransform(const float* lines, Index n, Index W, Index H, int dist) {
    Img img(H, W, std::numeric_limits<float>::max());
    drawLines(img, lines, n, 0.f);
    if (dist == L1) {
        colPassL1(img);
        img = transposed(img);
        colPassL1(img);
        return transposed(img);
    }
    colPassL2(img);
    img = transposed(img);
    colPassL2(img);
    img = transposed(img);
    if (dist == L2)
        for (auto& v : img.d) v = std::sqrt(v);
    return img;
}

// imgproc.h:38-84
static void lineIntegral(Img& img, float lineAngle) {
    float rx, ry;
    rasterizeVector(std::cos(lineAngle), std::sin(lineAngle), rx, ry);
    Index p0x = 0, p0y = 0;
    if (rx < 0) p0x += img.cols - 1;
    if (ry < 0) p0y += img.rows - 1;
    if (std::fabs(rx) == 1) {
        Index prev_x = p0x;
        for (Index i = 1; i < img.cols; ++i) {
            const Index px = p0x + i * Index(rx);
            const Index dy = static_cast<Index>(std::round(i * ry)) - static_cast<Index>(std::round((i - 1) * ry));
            const Index y1 = std::max(dy, Index(0));
            const Index y2 = std::max(-dy, Index(0));
            const Index len = img.rows - std::labs(dy);
            float* dst = &img.d[(size_t)px * img.rows + y1];
            const float* src = &img.d[(size_t)prev_x * img.rows + y2];
            for (Index j = 0; j < len; ++j) dst[j] += src[j];
            prev_x = px;
        }
    } else if (std::fabs(ry) == 1) {
        // row-major copy (imgproc.h:66): element (y,x) at r[y*cols+x]
        const Index Wc = img.cols, Hr = img.rows;
        std::vector<float> r((size_t)Wc * Hr);
        for (Index x = 0; x < Wc; ++x)
            for (Index y = 0; y < Hr; ++y) r[(size_t)y * Wc + x] = img.d[(size_t)x * Hr + y];
        Index prev_y = p0y;
        for (Index i = 1; i < Hr; ++i) {
            const Index dx = static_cast<Index>(std::round(i * rx)) - static_cast<Index>(std::round((i - 1) * rx));
            const Index py = p0y + i * Index(ry);
            const Index x1 = std::max(dx, Index(0));
            const Index x2 = std::max(-dx, Index(0));
            const Index len = Wc - std::labs(dx);
            float* dst = &r[(size_t)py * Wc + x1];
            const float* src = &r[(size_t)prev_y * Wc + x2];
            for (Index j = 0; j < len; ++j) dst[j] += src[j];
            prev_y = py;
        }
        for (Index x = 0; x < Wc; ++x)
            for (Index y = 0; y < Hr; ++y) img.d[(size_t)x * Hr + y] = r[(size_t)y * Wc + x];
    }
}

// ---------------------------------------------------------------------------------------------
// matching/featuremaps/dt3cpu.{h,cpp}
// ---------------------------------------------------------------------------------------------
// dt3cpu.h:93-114 over a sorted key array (std::map order); returns the index of the chosen key.
static Index closestOrientation(const float* keys, Index nkeys, const float* line) {
    const float line_angle = std::atan((line[3] - line[1]) / (line[2] - line[0]));   // math.h:295-299
    const Index it = std::upper_bound(keys, keys + nkeys, line_angle) - keys;
    if (it != nkeys && it != 0) {
        const float ud = std::abs(line_angle - keys[it]);
        const float ld = std::abs(line_angle - keys[it - 1]);
        if (ld < ud) return it - 1;
        return it;
    }
    const Index last = nkeys - 1;
    const float angle1 = line_angle - keys[0];
    const float angle2 = line_angle - keys[last];
    if (std::min(angle1, std::abs(angle1 - kPIf)) < std::min(angle2, std::abs(angle2 - kPIf))) return 0;
    return last;
}

// dt3cpu.cpp:109-116
static void getSceneCenteredTranslation(const float* scene, Index n, float padding, float shift[2], size_t size[2]) {
    float mn[2] = {scene[0], scene[1]}, mx[2] = {scene[0], scene[1]};
    for (Index i = 0; i < 2 * n; ++i)
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
        size[a] = (size_t)std::ceil(req + 1.f);
    }
}

// dt3cpu.cpp:30-75
static void minmaxTranslation(const float* tmpl, Index n, const float vec[2], const size_t fsize[2],
                              const float extra[2], float out[2]) {
    if (std::fabs(vec[0] - 0.f) <= 1e-5f + 0.f * std::fabs(0.f) &&
        std::fabs(vec[1] - 0.f) <= 1e-5f + 0.f * std::fabs(0.f)) {   // allClose(align_vec, {0,0})
        out[0] = kInf;
        out[1] = kInf;
        return;
    }
    const float size[2] = {(float)fsize[0], (float)fsize[1]};
    float mn[2] = {0, 0}, mx[2] = {0, 0};
    if (n > 0) {
        mn[0] = mx[0] = tmpl[0];
        mn[1] = mx[1] = tmpl[1];
    }
    for (Index i = 0; i < 2 * n; ++i)
        for (int a = 0; a < 2; ++a) {
            const float v = tmpl[2 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    for (int a = 0; a < 2; ++a) {
        mn[a] += extra[a];
        mx[a] += extra[a];
    }
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if ((size[0] - 1 - mx[0]) < 0 || (size[1] - 1 - mx[1]) < 0) { out[0] = out[1] = nan; return; }
    if (mn[0] < 0 || mn[1] < 0) { out[0] = out[1] = nan; return; }
    float m[2][4];
    for (int a = 0; a < 2; ++a) {
        m[a][0] = -mx[a];
        m[a][1] = -mn[a];
        m[a][2] = (size[a] - mx[a] - 1.f);
        m[a][3] = (size[a] - mn[a] - 1.f);
        for (int j = 0; j < 4; ++j) m[a][j] /= vec[a];
    }
    float E[2][2];   // E[0][a] = max of negative coeffs, E[1][a] = min of positive coeffs (NaN-propagating)
    for (int a = 0; a < 2; ++a) {
        float neg_max = 0, pos_min = 0;
        bool neg_nan = false, pos_nan = false;
        for (int j = 0; j < 4; ++j) {
            const bool sg = std::signbit(m[a][j]);
            const float pos = sg ? kInf : m[a][j];
            const float neg = sg ? m[a][j] : -kInf;
            if (std::isnan(pos)) pos_nan = true;
            if (std::isnan(neg)) neg_nan = true;
            if (j == 0) { neg_max = neg; pos_min = pos; }
            else {
                if (neg > neg_max) neg_max = neg;
                if (pos < pos_min) pos_min = pos;
            }
        }
        E[0][a] = neg_nan ? nan : neg_max;
        E[1][a] = pos_nan ? nan : pos_min;
    }
    const bool fin0 = std::isfinite(E[0][0]) && std::isfinite(E[1][0]);
    const bool fin1 = std::isfinite(E[0][1]) && std::isfinite(E[1][1]);
    if (fin0 && fin1) {
        out[0] = std::max(E[0][0], E[0][1]);
        out[1] = std::min(E[1][0], E[1][1]);
    } else if (fin0) {
        out[0] = E[0][0];
        out[1] = E[1][0];
    } else {
        out[0] = E[0][1];
        out[1] = E[1][1];
    }
}

// dt3cpu.cpp:77-107 on an ordered list of (key, plane)
static void propagateOrientation(std::vector<float>& keys, std::vector<Img>& planes, float coeff) {
    const int m = (int)planes.size();
    if (m == 0) return;
    const int fwd = static_cast<int>(std::ceil(1.5 * m));
    const int bwd = -static_cast<int>(std::floor(1.5 * m));
    auto propagate = [&](int start, int end, int step) {
        for (int c = start; c != end; c += step) {
            const int c1 = (m + ((c - step) % m)) % m;
            const int c2 = (m + (c % m)) % m;
            const float h = std::abs(keys[c1] - keys[c2]);
            const float min_h = std::min(h, std::abs(h - kPIf));
            const float w = coeff * min_h;
            float* a2 = planes[c2].d.data();
            const float* a1 = planes[c1].d.data();
            const size_t n = planes[c2].d.size();
            for (size_t i = 0; i < n; ++i) a2[i] = std::min(a2[i], a1[i] + w);
        }
    };
    propagate(0, fwd, 1);
    propagate(m, bwd, -1);
}

struct Dt3 {   // dt3cpu.h:46-63
    std::vector<float> keys;     // ascending (std::map order)
    std::vector<Img> planes;     // col-major H x W
    float shift[2]{0, 0};
    size_t size[2]{0, 0};        // (W, H)
};

// dt3cpu.h:174-234. stage: 0 = full; 1 = stop after the distance transform; 2 = stop after
// propagateOrientation (intermediate stages exist for stage-wise parity tests only).
static Dt3* buildCpuFeaturemap(const float* scene, Index n, size_t depth, float coeff, float padding, int dist,
                               int nthreads, int stage) {
    auto* fm = new Dt3();
    if (n == 0) return fm;
    getSceneCenteredTranslation(scene, n, padding, fm->shift, fm->size);
    std::vector<float> ts((size_t)4 * n);
    for (Index i = 0; i < 2 * n; ++i) {
        ts[2 * i] = scene[2 * i] + fm->shift[0];
        ts[2 * i + 1] = scene[2 * i + 1] + fm->shift[1];
    }
    std::map<float, int> keyset;   // std::set<float> (dt3cpu.h:188-190)
    for (size_t i = 0; i < depth; ++i) keyset[float(i) * kPIf / float(depth) - kPI_2f] = 0;
    for (auto& kv : keyset) fm->keys.push_back(kv.first);
    const Index D = (Index)fm->keys.size();
    std::vector<std::vector<float>> cls(D);
    for (Index i = 0; i < n; ++i) {
        const Index b = closestOrientation(fm->keys.data(), D, &ts[4 * i]);
        cls[b].insert(cls[b].end(), &ts[4 * i], &ts[4 * i] + 4);
    }
    fm->planes.resize(D);
    parallel_for((size_t)D, nthreads, [&](size_t a) {
        fm->planes[a] = distanceTransform(cls[a].data(), (Index)cls[a].size() / 4, (Index)fm->size[0], (Index)fm->size[1], dist);
    });
    if (stage == 1) return fm;
    propagateOrientation(fm->keys, fm->planes, coeff);
    if (stage == 2) return fm;
    for (Index a = 0; a < D; ++a) lineIntegral(fm->planes[a], fm->keys[a]);
    return fm;
}

static std::atomic<long long> g_evaluations{0}, g_lookups{0};

// dt3cpu.cpp:126-179 for one template and K translations
static void evaluate(const Dt3& fm, const float* tmpl, Index L, const float* transl, Index K, float* scores) {
    const Index D = (Index)fm.keys.size();
    std::vector<Index> bins(L);
    for (Index i = 0; i < L; ++i) bins[i] = closestOrientation(fm.keys.data(), D, tmpl + 4 * i);
    std::vector<float> per_line(L);
    for (Index k = 0; k < K; ++k) {
        const float ox = fm.shift[0] + transl[2 * k], oy = fm.shift[1] + transl[2 * k + 1];
        for (Index i = 0; i < L; ++i) {
            const int x1 = (int)(tmpl[4 * i] + ox), y1 = (int)(tmpl[4 * i + 1] + oy);
            const int x2 = (int)(tmpl[4 * i + 2] + ox), y2 = (int)(tmpl[4 * i + 3] + oy);
            const Img& f = fm.planes[bins[i]];
            const float a = f.at(y1, x1), b = f.at(y2, x2);
            per_line[i] = std::abs(a - b);
        }
        scores[k] = eigen_sum(per_line.data(), L);
    }
    g_evaluations.fetch_add(K, std::memory_order_relaxed);
    g_lookups.fetch_add(2 * K * L, std::memory_order_relaxed);
}

// ---------------------------------------------------------------------------------------------
// matching/searchstrategies/defaultsearch.{h,cpp}
// ---------------------------------------------------------------------------------------------
static std::vector<long> argsort_desc(const std::vector<float>& v) {   // math.h:107-116 with std::greater
    std::vector<long> ind(v.size());
    for (size_t i = 0; i < v.size(); ++i) ind[i] = (long)i;
    std::sort(ind.begin(), ind.end(), [&v](long const i1, long const i2) { return v[i1] > v[i2]; });
    return ind;
}

static void getCenteredRange(size_t center, size_t vec_size, size_t max_length, size_t& b, size_t& e) {
    b = std::max(0, int(center) - int(max_length / 2));   // defaultsearch.h:40-47
    e = std::min(size_t(b + max_length), vec_size);
    b = (size_t)std::max(0, int(e) - int(max_length));
}

struct Combo { long tmplLine, sceneLine; };

// ConcentricRangeStrategy (searchstrategies/concentricrange.h:35-60): optional radius filter on the scene lines
struct ConcFilter { bool on; float cx, cy, lo, hi; };

// concentricrange.h:73-84 filterInRange
static std::vector<long> filterInRange(const float* lines, Index n, float cx, float cy, float min_radius, float max_radius) {
    std::vector<long> idx;
    for (Index i = 0; i < n; ++i) {
        const float* l = lines + 4 * i;
        const float mx = (l[2] + l[0]) / 2 - cx, my = (l[3] + l[1]) / 2 - cy;   // getCenter - center_position
        const float rad = std::sqrt(mx * mx + my * my);
        if (rad > (min_radius - std::numeric_limits<float>::epsilon()) && rad < max_radius) idx.push_back((long)i);
    }
    return idx;
}

// defaultsearch.cpp:29-49
static std::vector<Combo> establishSearchStrategy(size_t maxT, size_t maxS, const float* tmpl, Index L,
                                                  const float* scene_all, Index M_all, const ConcFilter* cf = nullptr) {
    // concentricrange.cpp:29-60: filter, then the DefaultSearch procedure on the filtered lines, indices mapped back
    std::vector<long> filt;
    std::vector<float> fscene;
    const float* scene = scene_all;
    Index M = M_all;
    if (cf && cf->on) {
        filt = filterInRange(scene_all, M_all, cf->cx, cf->cy, cf->lo, cf->hi);
        if (filt.empty()) return {};
        for (long i : filt) fscene.insert(fscene.end(), scene_all + 4 * i, scene_all + 4 * i + 4);
        scene = fscene.data();
        M = (Index)filt.size();
    }
    std::vector<float> sl(M), tl(L);
    for (Index i = 0; i < M; ++i) sl[i] = line_length(scene + 4 * i);
    for (Index i = 0; i < L; ++i) tl[i] = line_length(tmpl + 4 * i);
    std::vector<long> ss = argsort_desc(sl);
    const std::vector<long> st = argsort_desc(tl);
    std::vector<float> sorted_len(M);
    for (Index i = 0; i < M; ++i) sorted_len[i] = sl[ss[i]];
    if (cf && cf->on)
        for (auto& v : ss) v = filt[(size_t)v];   // initial_sorted_filtered_scene_idx
    std::vector<Combo> out;
    const size_t nt = std::min((size_t)L, maxT);
    for (size_t r = 0; r < nt; ++r) {
        const long ti = st[r];
        const float value = tl[ti];
        // math.h:138-146 binarySearch with std::greater
        auto it = std::lower_bound(sorted_len.begin(), sorted_len.end(), value, std::greater<float>());
        size_t closest;
        if (it == sorted_len.begin()) closest = 0;
        else if (it == sorted_len.end()) closest = (size_t)std::distance(sorted_len.begin(), it - 1);
        else closest = std::abs(value - *it) < std::abs(value - *(it - 1)) ? (size_t)std::distance(sorted_len.begin(), it)
                                                                          : (size_t)std::distance(sorted_len.begin(), it - 1);
        size_t b, e;
        getCenteredRange(closest, (size_t)M, maxS, b, e);
        for (size_t i = b; i < e; ++i) out.push_back(Combo{ti, ss.at(i)});
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// matching/optimizestrategies/{batch,default}optimize.cpp
// ---------------------------------------------------------------------------------------------
struct OptResult { bool has; float score, tx, ty; };

// batchoptimize.cpp:15-99 (batch >= 1) / defaultoptimize.cpp:15-64 (batch == 0)
static OptResult optimize_one(const Dt3& fm, const float* tmpl, Index L, float avx, float avy, long batch) {
    OptResult none{false, 0, 0, 0};
    if (relativelyEqual(std::fabs(avx) + std::fabs(avy), 0.f)) return none;
    float sv[2];
    rasterizeVector(avx, avy, sv[0], sv[1]);
    float mm[2];
    minmaxTranslation(tmpl, L, sv, fm.size, fm.shift, mm);
    const float min_mul = mm[0], max_mul = mm[1];
    if (!std::isfinite(min_mul) || !std::isfinite(max_mul)) return none;
    const float zero[2] = {0.f, 0.f};
    float initial;
    evaluate(fm, tmpl, L, zero, 1, &initial);
    std::vector<float> tr{0.f, 0.f};
    std::vector<float> scores{initial};
    if (batch <= 0) {   // DefaultOptimize
        for (long mul = 1; mul <= static_cast<long>(max_mul); ++mul) {
            const float t[2] = {(float)mul * sv[0], (float)mul * sv[1]};
            float s;
            evaluate(fm, tmpl, L, t, 1, &s);
            if (s > scores.back()) break;
            tr.push_back(t[0]); tr.push_back(t[1]);
            scores.push_back(s);
        }
        for (long mul = -1; mul >= static_cast<long>(min_mul); --mul) {
            const float t[2] = {(float)mul * sv[0], (float)mul * sv[1]};
            float s;
            evaluate(fm, tmpl, L, t, 1, &s);
            if (s > scores.back()) break;
            tr.push_back(t[0]); tr.push_back(t[1]);
            scores.push_back(s);
        }
    } else {
        std::vector<float> bt, bs;
        for (long mul = 1; mul <= static_cast<long>(max_mul); mul += batch) {
            bt.clear();
            for (long b = mul; b < mul + batch && b <= static_cast<long>(max_mul); ++b) {
                bt.push_back((float)b * sv[0]);
                bt.push_back((float)b * sv[1]);
            }
            bs.resize(bt.size() / 2);
            evaluate(fm, tmpl, L, bt.data(), (Index)bs.size(), bs.data());
            const int am = (int)std::distance(bs.begin(), std::min_element(bs.begin(), bs.end()));
            if (bs[am] > scores.back()) break;
            tr.push_back(bt[2 * am]); tr.push_back(bt[2 * am + 1]);
            scores.push_back(bs[am]);
            if (bs[am] < bs.back()) break;
        }
        for (long mul = -1; mul >= static_cast<long>(min_mul); mul -= batch) {
            bt.clear();
            for (long b = mul; b > mul - batch && b >= static_cast<long>(min_mul); --b) {
                bt.push_back((float)b * sv[0]);
                bt.push_back((float)b * sv[1]);
            }
            bs.resize(bt.size() / 2);
            evaluate(fm, tmpl, L, bt.data(), (Index)bs.size(), bs.data());
            const int am = (int)std::distance(bs.begin(), std::min_element(bs.begin(), bs.end()));
            if (bs[am] > scores.back()) break;
            tr.push_back(bt[2 * am]); tr.push_back(bt[2 * am + 1]);
            scores.push_back(bs[am]);
            if (bs[am] < bs.back()) break;
        }
    }
    const size_t best = (size_t)std::distance(scores.begin(), std::min_element(scores.begin(), scores.end()));
    return OptResult{true, scores[best], tr[2 * best], tr[2 * best + 1]};
}

struct MatchRec { int32_t tmpl_idx; float score; float t[6]; };   // t = row-major 2x3

// defaultmatch.cpp:32-89
static std::vector<MatchRec> search(const Dt3& fm, const float* tl, const int32_t* off, Index T, const float* scene,
                                    Index M, size_t maxT, size_t maxS, long batch, int nthreads,
                                    std::vector<int32_t>* hyp_out /* optional: (tmpl, tmplLine, sceneLine, rev) */,
                                    const ConcFilter* cf = nullptr) {
    std::vector<MatchRec> all;
    if (T == 0 || M == 0 || (fm.size[0] == 0 && fm.size[1] == 0)) return all;
    std::vector<std::vector<float>> aligned;
    std::vector<int> tidx;
    std::vector<float> avec;
    std::vector<Mat23> tf;
    for (Index t = 0; t < T; ++t) {
        const float* tmpl = tl + 4 * (size_t)off[t];
        const Index L = off[t + 1] - off[t];
        if (L == 0) continue;
        for (const Combo& cb : establishSearchStrategy(maxT, maxS, tmpl, L, scene, M, cf)) {
            const float* sline = scene + 4 * cb.sceneLine;
            const float* tline = tmpl + 4 * cb.tmplLine;
            float ax, ay;
            unit_vec(sline, ax, ay);
            Mat23 T1, T2;
            align(tline, sline, T1, T2);
            for (int rev = 0; rev < 2; ++rev) {
                const Mat23& Tr = rev ? T2 : T1;
                tf.push_back(Tr);
                tidx.push_back((int)t);
                std::vector<float> a((size_t)4 * L);
                for (Index p = 0; p < 2 * L; ++p) transform_pt(Tr, tmpl[2 * p], tmpl[2 * p + 1], a[2 * p], a[2 * p + 1]);
                aligned.push_back(std::move(a));
                avec.push_back(ax);
                avec.push_back(ay);
                if (hyp_out) {
                    hyp_out->push_back((int32_t)t);
                    hyp_out->push_back((int32_t)cb.tmplLine);
                    hyp_out->push_back((int32_t)cb.sceneLine);
                    hyp_out->push_back(rev);
                }
            }
        }
    }
    std::vector<OptResult> res(aligned.size());
    parallel_for(aligned.size(), nthreads, [&](size_t h) {
        res[h] = optimize_one(fm, aligned[h].data(), (Index)aligned[h].size() / 4, avec[2 * h], avec[2 * h + 1], batch);
    });
    for (size_t h = 0; h < aligned.size(); ++h) {
        if (!res[h].has) continue;
        const Mat23& Tr = tf[h];   // combine(translation, transform) math.h:427-432
        all.push_back(MatchRec{tidx[h], res[h].score, {Tr.r00, Tr.r01, Tr.tx + res[h].tx, Tr.r10, Tr.r11, Tr.ty + res[h].ty}});
    }
    return all;
}

}   // namespace orc

// =============================================================================================
// C interface (ctypes)
// =============================================================================================
extern "C" {

void orc_rasterize_vector(const float* v, float* out) { orc::rasterizeVector(v[0], v[1], out[0], out[1]); }

int orc_rasterize_line(const float* line, long* out_xy, int cap) {
    std::vector<orc::Index> xs, ys;
    orc::rasterizeLine(line, xs, ys);
    for (size_t i = 0; i < xs.size() && (int)i < cap; ++i) {
        out_xy[2 * i] = xs[i];
        out_xy[2 * i + 1] = ys[i];
    }
    return (int)xs.size();
}

// returns number of kept lines (deleteOob=1) or n (deleteOob=0, purged lines zeroed)
int orc_clip_lines(const float* lines, int n, float xmin, float xmax, float ymin, float ymax, int deleteOob, float* out) {
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        float l[4] = {lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]};
        const bool ok = orc::clip_line(l, xmin, xmax, ymin, ymax);
        if (deleteOob) {
            if (ok) { std::memcpy(out + 4 * kept, l, sizeof l); ++kept; }
        } else {
            if (!ok) l[0] = l[1] = l[2] = l[3] = 0.f;
            std::memcpy(out + 4 * i, l, sizeof l);
            ++kept;
        }
    }
    return kept;
}

// img: row-major H x W in/out
void orc_draw_lines(float* img, int H, int W, const float* lines, int n, float color) {
    orc::Img im(H, W, 0.f);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) im.at(y, x) = img[(size_t)y * W + x];
    orc::drawLines(im, lines, n, color);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) img[(size_t)y * W + x] = im.at(y, x);
}

void orc_distance_transform(const float* lines, int n, int W, int H, int dist, float* out) {
    orc::Img im = orc::distanceTransform(lines, n, W, H, dist);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = im.at(y, x);
}

// literal 1-D pass of imgproc.h:91-130 on one scan-line (in place)
void orc_dt_pass_l2_1d(float* f, int n) {
    orc::Img im(n, 1, 0.f);
    std::memcpy(im.d.data(), f, sizeof(float) * n);
    orc::colPassL2(im);
    std::memcpy(f, im.d.data(), sizeof(float) * n);
}

void orc_line_integral(float* img, int H, int W, float angle) {
    orc::Img im(H, W, 0.f);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) im.at(y, x) = img[(size_t)y * W + x];
    orc::lineIntegral(im, angle);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) img[(size_t)y * W + x] = im.at(y, x);
}

void orc_scene_shift(const float* scene, int n, float padding, float* shift, uint64_t* size) {
    size_t s[2];
    orc::getSceneCenteredTranslation(scene, n, padding, shift, s);
    size[0] = s[0];
    size[1] = s[1];
}

int orc_closest_orientation(const float* keys, int nkeys, const float* line) {
    return (int)orc::closestOrientation(keys, nkeys, line);
}

void orc_angle_keys(int depth, float* keys) {
    for (int i = 0; i < depth; ++i) keys[i] = float(i) * orc::kPIf / float(depth) - orc::kPI_2f;
}

// planes: D x H x W row-major, in place
void orc_propagate_orientation(float* planes, int D, int H, int W, const float* keys, float coeff) {
    std::vector<float> k(keys, keys + D);
    std::vector<orc::Img> p(D);
    const size_t n = (size_t)H * W;
    for (int d = 0; d < D; ++d) {
        p[d] = orc::Img(H, W, 0.f);
        std::memcpy(p[d].d.data(), planes + d * n, n * sizeof(float));   // layout-agnostic elementwise op
    }
    orc::propagateOrientation(k, p, coeff);
    for (int d = 0; d < D; ++d) std::memcpy(planes + d * n, p[d].d.data(), n * sizeof(float));
}

void orc_minmax_translation(const float* tmpl, int n, const float* vec, uint64_t W, uint64_t H, const float* extra, float* out) {
    const size_t fs[2] = {(size_t)W, (size_t)H};
    orc::minmaxTranslation(tmpl, n, vec, fs, extra, out);
}

float orc_eigen_sum(const float* c, int n) { return orc::eigen_sum(c, n); }

void orc_align(const float* tmpl_line, const float* scene_line, float* t1, float* t2) {
    orc::Mat23 a, b;
    orc::align(tmpl_line, scene_line, a, b);
    std::memcpy(t1, &a, sizeof a);
    std::memcpy(t2, &b, sizeof b);
}

void orc_transform(const float* lines, int n, const float* t, float* out) {
    orc::Mat23 T;
    std::memcpy(&T, t, sizeof T);
    for (int p = 0; p < 2 * n; ++p) orc::transform_pt(T, lines[2 * p], lines[2 * p + 1], out[2 * p], out[2 * p + 1]);
}

// ---- feature map object ----
void* orc_dt3_build(const float* scene, int n, int depth, float coeff, float padding, int dist, int nthreads, int stage) {
    return orc::buildCpuFeaturemap(scene, n, (size_t)depth, coeff, padding, dist, nthreads, stage);
}
void orc_dt3_free(void* h) { delete (orc::Dt3*)h; }
void orc_dt3_info(const void* h, int* depth, uint64_t* W, uint64_t* H, float* shift) {
    const auto* fm = (const orc::Dt3*)h;
    *depth = (int)fm->keys.size();
    *W = fm->size[0];
    *H = fm->size[1];
    shift[0] = fm->shift[0];
    shift[1] = fm->shift[1];
}
void orc_dt3_keys(const void* h, float* keys) {
    const auto* fm = (const orc::Dt3*)h;
    std::memcpy(keys, fm->keys.data(), fm->keys.size() * sizeof(float));
}
// row-major H x W copy of plane d
void orc_dt3_plane(const void* h, int d, float* out) {
    const auto* fm = (const orc::Dt3*)h;
    const orc::Img& im = fm->planes[d];
    const orc::Index B = 32;
    for (orc::Index y0 = 0; y0 < im.rows; y0 += B)
        for (orc::Index x0 = 0; x0 < im.cols; x0 += B)
            for (orc::Index y = y0; y < std::min(im.rows, y0 + B); ++y)
                for (orc::Index x = x0; x < std::min(im.cols, x0 + B); ++x) out[(size_t)y * im.cols + x] = im.at(y, x);
}
void orc_dt3_minmax_translation(const void* h, const float* tmpl, int n, const float* vec, float* out) {
    const auto* fm = (const orc::Dt3*)h;
    orc::minmaxTranslation(tmpl, n, vec, fm->size, fm->shift, out);
}
void orc_dt3_evaluate(const void* h, const float* tmpl, int L, const float* transl, int K, float* scores) {
    orc::evaluate(*(const orc::Dt3*)h, tmpl, L, transl, K, scores);
}

// ---- search ----
int orc_default_search(const float* tmpl, int L, const float* scene, int M, uint64_t maxT, uint64_t maxS, long* out_pairs, int cap) {
    auto v = orc::establishSearchStrategy((size_t)maxT, (size_t)maxS, tmpl, L, scene, M);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) {
        out_pairs[2 * i] = v[i].tmplLine;
        out_pairs[2 * i + 1] = v[i].sceneLine;
    }
    return (int)v.size();
}

int orc_filter_in_range(const float* lines, int n, float cx, float cy, float lo, float hi, long* out) {
    auto v = orc::filterInRange(lines, n, cx, cy, lo, hi);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}

int orc_concentric_search(const float* tmpl, int L, const float* scene, int M, uint64_t maxT, uint64_t maxS, float cx, float cy,
                          float lo, float hi, long* out_pairs, int cap) {
    const orc::ConcFilter cf{true, cx, cy, lo, hi};
    auto v = orc::establishSearchStrategy((size_t)maxT, (size_t)maxS, tmpl, L, scene, M, &cf);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) {
        out_pairs[2 * i] = v[i].tmplLine;
        out_pairs[2 * i + 1] = v[i].sceneLine;
    }
    return (int)v.size();
}

void orc_centered_range(uint64_t c, uint64_t n, uint64_t len, uint64_t* b, uint64_t* e) {
    size_t bb, ee;
    orc::getCenteredRange((size_t)c, (size_t)n, (size_t)len, bb, ee);
    *b = bb;
    *e = ee;
}

// one hypothesis; batch<=0 → DefaultOptimize. out = {score, tx, ty}; returns has_value
int orc_optimize_one(const void* h, const float* tmpl, int L, const float* align_vec, long batch, float* out) {
    orc::OptResult r = orc::optimize_one(*(const orc::Dt3*)h, tmpl, L, align_vec[0], align_vec[1], batch);
    out[0] = r.score;
    out[1] = r.tx;
    out[2] = r.ty;
    return r.has ? 1 : 0;
}

// Full DefaultMatch search. Returns number of matches (may exceed cap; only cap are written).
// hyp (optional, 4 ints per hypothesis, capacity hyp_cap hypotheses) receives the hypothesis list.
long orc_search(const void* h, const float* tmpl_lines, const int32_t* tmpl_offsets, int n_tmpl, const float* scene,
                int n_scene, uint64_t maxT, uint64_t maxS, long batch, int nthreads, void* out_matches, long cap,
                int32_t* hyp, long hyp_cap, long* n_hyp, const float* concentric /* null or {cx, cy, lo, hi} */) {
    std::vector<int32_t> hv;
    orc::ConcFilter cf{false, 0, 0, 0, 0};
    if (concentric) cf = orc::ConcFilter{true, concentric[0], concentric[1], concentric[2], concentric[3]};
    auto m = orc::search(*(const orc::Dt3*)h, tmpl_lines, tmpl_offsets, n_tmpl, scene, n_scene, (size_t)maxT, (size_t)maxS,
                         batch, nthreads, (hyp || n_hyp) ? &hv : nullptr, concentric ? &cf : nullptr);
    if (n_hyp) *n_hyp = (long)hv.size() / 4;
    if (hyp) std::memcpy(hyp, hv.data(), sizeof(int32_t) * std::min<size_t>(hv.size(), (size_t)hyp_cap * 4));
    const long n = (long)m.size();
    if (out_matches && cap > 0) std::memcpy(out_matches, m.data(), sizeof(orc::MatchRec) * (size_t)std::min(n, cap));
    return n;
}

void orc_template_lengths(const float* tmpl_lines, const int32_t* off, int T, float* out) {   // math.h:319-324
    for (int t = 0; t < T; ++t) {
        const int L = off[t + 1] - off[t];
        std::vector<float> len(L);
        for (int i = 0; i < L; ++i) len[i] = orc::line_length(tmpl_lines + 4 * ((size_t)off[t] + i));
        out[t] = orc::eigen_sum(len.data(), L);
    }
}

// kind 0 = DefaultPenalty (defaultpenalty.cpp:35-41), 1 = ExponentialPenalty(tau) (exponentialpenalty.cpp:39-46).
// returns 0, or -1 for the std::out_of_range case.
int orc_penalize(int kind, float tau, void* matches, long n, const float* lengths, long n_lengths) {
    auto* m = (orc::MatchRec*)matches;
    for (long i = 0; i < n; ++i) {
        const size_t idx = (size_t)m[i].tmpl_idx;
        if (idx >= (size_t)n_lengths) return -1;
        const float len = std::max(lengths[idx], 1e-6f);
        m[i].score = kind == 0 ? m[i].score / len : m[i].score / std::pow(len, tau);
    }
    return 0;
}

void orc_sort_matches(void* matches, long n) {   // python/src/matching.cpp:302-307
    auto* m = (orc::MatchRec*)matches;
    std::sort(m, m + n, [](const orc::MatchRec& a, const orc::MatchRec& b) { return a.score < b.score; });
}

void orc_stats_reset() { orc::g_evaluations = 0; orc::g_lookups = 0; }
void orc_stats_get(long long* evaluations, long long* lookups) {
    *evaluations = orc::g_evaluations.load();
    *lookups = orc::g_lookups.load();
}
int orc_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}   // extern "C"
