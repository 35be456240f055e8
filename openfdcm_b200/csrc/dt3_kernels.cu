// dt3_kernels.cu — hot path 1: DT3 feature-map build kernels (sm_100a).
//
// Pipeline (reference: buildCpuFeaturemap, matching/featuremaps/dt3cpu.h:174-234):
//   K1 raster_kernel            drawLines: clipLines + rasterizeLine per orientation plane  -> 1-bit edge mask
//   K2 dt_col_exact_kernel      first (vertical) pass of distanceTransform on the binary mask -> u16 distance
//      dt_pass_literal_kernel   Felzenszwalb lower-envelope pass, literal incl. the in-place aliasing
//      dt_row_l1_kernel         L1 second pass
//   K3 propagate_kernel         propagateOrientation: 4*D circular min-plus steps, D values in registers
//   K4 integral_kernel          lineIntegral: sequential fp32 running sums along each plane's discrete lines
// Layout: orientation-major [D][H][pitch] fp32, pitch % 32 == 0.
// All float arithmetic is non-fused (-fmad=false) and ordered exactly as the reference's expressions.
#include "common.cuh"
#include "kernels.h"

namespace fdcm {

// =============================================================================================
// K1: clip + rasterise scene lines into the per-plane edge bit mask
// =============================================================================================
__device__ __forceinline__ int outcode_dev(float x, float y, float xmax, float ymax) {
    int code = 0;   // core/src/drawing.cpp:37-50 (LEFT=1 RIGHT=2 BOTTOM=4 TOP=8), box [0,xmax]x[0,ymax]
    if (x < 0.f) code |= 1; else if (x > xmax) code |= 2;
    if (y < 0.f) code |= 4; else if (y > ymax) code |= 8;
    return code;
}

// Cohen-Sutherland (core/src/drawing.cpp:64-112); returns false when the line is purged
__device__ bool clip_line_dev(float& x1, float& y1, float& x2, float& y2, float xmax, float ymax) {
    int c1 = outcode_dev(x1, y1, xmax, ymax);
    int c2 = outcode_dev(x2, y2, xmax, ymax);
    for (int it = 0; it < 16; ++it) {   // the reference loop ends after <= 4 clips; bound guards NaN input
        if (c1 == 0 && c2 == 0) return true;
        if (c1 & c2) return false;
        if (c1 != 0) {
            if (c1 & 8)      { x1 = x1 + (x2 - x1) * (ymax - y1) / (y2 - y1); y1 = ymax; }
            else if (c1 & 4) { x1 = x1 + (x2 - x1) * (0.f - y1) / (y2 - y1); y1 = 0.f; }
            else if (c1 & 2) { y1 = y1 + (y2 - y1) * (xmax - x1) / (x2 - x1); x1 = xmax; }
            else if (c1 & 1) { y1 = y1 + (y2 - y1) * (0.f - x1) / (x2 - x1); x1 = 0.f; }
            c1 = outcode_dev(x1, y1, xmax, ymax);
            continue;
        }
        if (c2 & 8)      { x2 = x2 + (x1 - x2) * (ymax - y2) / (y1 - y2); y2 = ymax; }
        else if (c2 & 4) { x2 = x2 + (x1 - x2) * (0.f - y2) / (y1 - y2); y2 = 0.f; }
        else if (c2 & 2) { y2 = y2 + (y1 - y2) * (xmax - x2) / (x1 - x2); x2 = xmax; }
        else if (c2 & 1) { y2 = y2 + (y1 - y2) * (0.f - x2) / (x1 - x2); x2 = 0.f; }
        c2 = outcode_dev(x2, y2, xmax, ymax);
    }
    return false;
}

// relativelyEqual(x, 0.0f) (core/math.h:183-189): |x| <= eps_f + 1e-10*|x| evaluated in double
__device__ __forceinline__ bool rel_eq_zero(float x) {
    const double ax = (double)fabsf(x);
    return ax <= (double)FLT_EPSILON + 1e-10 * ax;
}

// Eigen 3.4.0 LinSpaced<float>(n, lo, hi)(i) (see oracle / SURVEY App. A.7)
struct LinSpacedDev {
    float lo, hi, step;
    int size1;
    bool flip;
    __device__ LinSpacedDev(int n, float lo_, float hi_)
        : lo(lo_), hi(hi_), step(n == 1 ? 0.f : (hi_ - lo_) / (float)(n - 1)), size1(n == 1 ? 1 : n - 1),
          flip(fabsf(hi_) < fabsf(lo_)) {}
    __device__ __forceinline__ float at(int i) const {
        if (flip) return (i == 0) ? lo : (hi - (float)(size1 - i) * step);
        return (i == size1) ? hi : (lo + (float)i * step);
    }
};

__device__ __forceinline__ void set_edge(uint32_t* mask, const MapDims& dm, int plane, long long x, long long y) {
    if (x < 0 || y < 0 || x >= dm.W || y >= dm.H) return;   // unreachable after clipping; guards NaN input
    atomicOr(mask + ((size_t)plane * dm.H + (size_t)y) * dm.wwords + (x >> 5), 1u << (x & 31));
}

// one warp per scene line (lines already shifted by the scene translation, dt3cpu.h:185)
__global__ void __launch_bounds__(256) raster_kernel(const float4* __restrict__ lines, const int32_t* __restrict__ bins,
                                                     int n_lines, MapDims dm, uint32_t* __restrict__ mask) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_lines) return;
    const float4 l = lines[warp];
    const int plane = bins[warp];
    float p1x = l.x, p1y = l.y, p2x = l.z, p2y = l.w;
    if (!clip_line_dev(p1x, p1y, p2x, p2y, (float)(dm.W - 1), (float)(dm.H - 1))) return;   // core/drawing.h:117-118
    // rasterizeLine (core/drawing.h:74-102)
    if (fabsf(p2x - p1x) <= 1e-5f && fabsf(p2y - p1y) <= 1e-5f) {
        if (lane == 0) set_edge(mask, dm, plane, round_to_ll(p1x), round_to_ll(p1y));
        return;
    }
    const float lvx = p2x - p1x, lvy = p2y - p1y;
    float rx, ry;
    rasterize_vector_dev(lvx, lvy, rx, ry);
    if (rel_eq_zero(rx)) {
        const int size = (int)(lvy / ry) + 1;
        const LinSpacedDev ly(size, p1y, p2y);
        const long long x = round_to_ll(p1x);
        for (int i = lane; i < size; i += 32) set_edge(mask, dm, plane, x, round_to_ll(ly.at(i)));
        return;
    }
    if (rel_eq_zero(ry)) {
        const int size = (int)(lvx / rx) + 1;
        const LinSpacedDev lx(size, p1x, p2x);
        const long long y = round_to_ll(p1y);
        for (int i = lane; i < size; i += 32) set_edge(mask, dm, plane, round_to_ll(lx.at(i)), y);
        return;
    }
    const float a = lvx / rx, b = lvy / ry;
    const int size = (int)((a < b) ? b : a) + 1;   // std::max(a, b)
    const LinSpacedDev lx(size, p1x, p2x), ly(size, p1y, p2y);
    for (int i = lane; i < size; i += 32) set_edge(mask, dm, plane, round_to_ll(lx.at(i)), round_to_ll(ly.at(i)));
}

// =============================================================================================
// K2a: vertical pass on the binary mask, exact regime.
// With a {0, FLT_MAX} input and 2*(side-1)^2 < 2^24 every quantity of the reference's first
// _distanceTransformColumnPassL2 call (core/imgproc.h:91-130) is an exactly representable integer,
// so its output is (distance to the nearest edge pixel in the column)^2; we store the distance
// itself (u16; 0xFFFF = no edge in the column = FLT_MAX).  The same array feeds the L1 transform.
// =============================================================================================
__global__ void __launch_bounds__(128) dt_col_exact_kernel(const uint32_t* __restrict__ mask, MapDims dm,
                                                           uint16_t* __restrict__ g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (x >= dm.W) return;
    const uint32_t* m = mask + (size_t)d * dm.H * dm.wwords + (x >> 5);
    const uint32_t bit = 1u << (x & 31);
    uint16_t* gp = g + (size_t)d * dm.plane_elems + x;
    int last = -1;
#pragma unroll 8
    for (int y = 0; y < dm.H; ++y) {
        if (m[(size_t)y * dm.wwords] & bit) last = y;
        gp[(size_t)y * dm.pitch] = last < 0 ? kNoEdge16 : (uint16_t)(y - last);
    }
    int next = -1;
#pragma unroll 8
    for (int y = dm.H - 1; y >= 0; --y) {
        if (m[(size_t)y * dm.wwords] & bit) next = y;
        const uint16_t dn = next < 0 ? kNoEdge16 : (uint16_t)(next - y);
        const uint16_t up = gp[(size_t)y * dm.pitch];
        gp[(size_t)y * dm.pitch] = up < dn ? up : dn;
    }
}

// general regime: materialise the {0, FLT_MAX} image (core/imgproc.h:174-175)
__global__ void __launch_bounds__(256) mask_to_float_kernel(const uint32_t* __restrict__ mask, MapDims dm,
                                                            float* __restrict__ planes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)dm.D * dm.plane_elems;
    if (i >= total) return;
    const int x = (int)(i % dm.pitch);
    const size_t row = i / dm.pitch;
    const uint32_t w = mask[row * dm.wwords + (x >> 5)];
    planes[i] = ((w >> (x & 31)) & 1u) ? 0.f : FLT_MAX;
}

// =============================================================================================
// K2b: one literal Felzenszwalb pass (core/imgproc.h:91-130) per scan-line, one thread per scan-line.
// Scan-line l of plane d starts at base + l*line_stride, element q at + q*elem_stride (rows:
// line_stride = pitch, elem_stride = 1; columns: the other way round).  The envelope stack (v, z)
// lives in a global workspace with the same indexing.  The second loop reproduces the reference's
// in-place read of img(v_k, i): when v_k < q the value has already been overwritten.
// kFromG: input is the u16 vertical distance of dt_col_exact_kernel (f = g*g, 0xFFFF -> FLT_MAX).
// =============================================================================================
struct EnvEntry { int v; float z; };

template <bool kFromG>
__global__ void __launch_bounds__(128) dt_pass_literal_kernel(const uint16_t* __restrict__ g, float* planes, MapDims dm,
                                                              EnvEntry* stack, int n, int n_lines, size_t elem_stride,
                                                              size_t line_stride) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (l >= n_lines) return;
    const size_t base = (size_t)d * dm.plane_elems + (size_t)l * line_stride;
    float* out = planes + base;
    const uint16_t* gin = g + base;
    EnvEntry* st = stack + base;
    auto F = [&](int q) -> float {
        if (kFromG) {
            const unsigned gv = gin[(size_t)q * elem_stride];
            return gv == kNoEdge16 ? FLT_MAX : (float)(gv * gv);
        }
        return out[(size_t)q * elem_stride];
    };
    // first loop: lower envelope (imgproc.h:101-121); top of stack cached in registers
    int k = 0, vk = 0;
    float fvk = F(0), zk = -INFINITY;
    st[0] = EnvEntry{0, -INFINITY};
    for (int q = 1; q < n; ++q) {
        const float fq = F(q);
        const float sqq = (float)((long long)q * q);
        while (true) {
            const float s = (fq + sqq - fvk - (float)((long long)vk * vk)) / (float)(2LL * q - 2LL * vk);
            if (s > zk) {
                ++k;
                st[(size_t)k * elem_stride] = EnvEntry{q, s};
                vk = q; fvk = fq; zk = s;
                break;
            }
            --k;
            const EnvEntry e = st[(size_t)k * elem_stride];
            vk = e.v; zk = e.z; fvk = F(vk);
        }
    }
    // second loop (imgproc.h:122-128), in place
    int k2 = 0, v = 0;
    float znext = (k >= 1) ? st[elem_stride].z : INFINITY;
    for (int q = 0; q < n; ++q) {
        while (znext < (float)q) {
            ++k2;
            v = st[(size_t)k2 * elem_stride].v;
            znext = (k2 + 1 <= k) ? st[(size_t)(k2 + 1) * elem_stride].z : INFINITY;
        }
        const long long dq = (long long)q - v;
        const float src = (kFromG && v >= q) ? F(v) : out[(size_t)v * elem_stride];
        out[(size_t)q * elem_stride] = src + (float)(dq * dq);
    }
}

// =============================================================================================
// K2b (exact regime): horizontal pass of the L2 / L2^2 transform, one warp per row, row in shared memory.
//
// In the exact regime (2*(side-1)^2 < 2^24) every quantity of the reference's second
// _distanceTransformColumnPassL2 call (core/imgproc.h:91-130) is an exactly representable integer, so
//   (1) the envelope it builds is the true lower envelope of the parabolas f[v] + (q-v)^2, f = g^2, and the
//       vertex that owns an integer q is the LEFTMOST argmin_v f[v] + (q-v)^2 (`while (z[k+1] < q)` keeps
//       the left parabola on a tie; FLT_MAX columns never own a pixel);
//   (2) its in-place second loop computes out[q] = (u < q ? out[u] : f[u]) + (q-u)^2 with u = owner(q),
//       a sum of integers, exact in any order.
// owner() is non-decreasing in q, so it is found by divide and conquer: the owner of the midpoint of an
// interval lies between the owners of its end points (equal end owners resolve the whole interval).
// Lanes take one query each; brackets longer than kCoopLen are scanned by the whole warp.  The chained
// values are then resolved in place, 32 pixels at a time, exactly like the reference's left-to-right sweep.
// =============================================================================================
constexpr uint32_t kBigF = 0x3FFFFFFFu;   // stands for FLT_MAX: never wins against a finite parabola
constexpr int kCoopLen = 40;

// leftmost argmin of f[v] + (q-v)^2 over v in [lo, hi], all 32 lanes cooperating
__device__ __forceinline__ int coop_owner(const uint32_t* f, int q, int lo, int hi, int lane) {
    unsigned long long best = ~0ull;
    for (int v = lo + lane; v <= hi; v += 32) {
        const int d = q - v;
        const unsigned long long key = ((unsigned long long)(f[v] + (uint32_t)(d * d)) << 16) | (unsigned)v;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    return (int)(best & 0xFFFFu);
}

__global__ void __launch_bounds__(128) dt_row_exact_kernel(const uint16_t* __restrict__ g, float* __restrict__ planes,
                                                           MapDims dm, int n_rows_total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    const int n = dm.W;
    uint32_t* f = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * dm.pitch;
    uint16_t* owner = reinterpret_cast<uint16_t*>(reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warps * dm.pitch) +
                      (size_t)warp * dm.pitch;
    const int row = blockIdx.x * warps + warp;   // row of the [D*H][pitch] stack of planes
    if (row >= n_rows_total) return;
    const uint16_t* gin = g + (size_t)row * dm.pitch;
    float* out = planes + (size_t)row * dm.pitch;

    // ---- load g, f = g^2; first / last finite column ----
    int cmin = 0x7fffffff, cmax = -1;
    for (int x = lane * 2; x < n; x += 64) {
        const uint32_t two = *reinterpret_cast<const uint32_t*>(gin + x);   // pitch is even: x+1 < pitch
        const uint32_t g0 = two & 0xFFFFu, g1 = two >> 16;
        f[x] = g0 == kNoEdge16 ? kBigF : g0 * g0;
        if (g0 != kNoEdge16) { cmin = min(cmin, x); cmax = max(cmax, x); }
        if (x + 1 < n) {
            f[x + 1] = g1 == kNoEdge16 ? kBigF : g1 * g1;
            if (g1 != kNoEdge16) { cmin = min(cmin, x + 1); cmax = max(cmax, x + 1); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    if (cmax < 0) {   // no edge pixel in this plane: the row stays FLT_MAX (imgproc.h:174)
        for (int x = lane; x < n; x += 32) out[x] = FLT_MAX;
        return;
    }
    __syncwarp();

    // ---- owners by divide and conquer ----
    int P = 1;
    while (P < n) P <<= 1;
    for (int s = P; s >= 1; s >>= 1) {
        // queries of this level: q = s-1 + 2s*j < n (q+1 = s * odd)
        if (s - 1 >= n) continue;
        const int nq = (n - (s - 1) + 2 * s - 1) / (2 * s);
        if (nq < 32) {
            for (int j = 0; j < nq; ++j) {
                const int q = s - 1 + 2 * s * j;
                const int lo = (q - s >= 0) ? owner[q - s] : cmin;
                const int hi = (q + s < n) ? owner[q + s] : cmax;
                const int o = (lo == hi) ? lo : coop_owner(f, q, lo, hi, lane);
                if (lane == 0) owner[q] = (uint16_t)o;
            }
        } else {
            for (int j0 = 0; j0 < nq; j0 += 32) {
                const int j = j0 + lane;
                const bool act = j < nq;
                const int q = s - 1 + 2 * s * j;
                int lo = 0, hi = 0;
                if (act) {
                    lo = (q - s >= 0) ? owner[q - s] : cmin;
                    hi = (q + s < n) ? owner[q + s] : cmax;
                }
                const bool is_long = act && (hi - lo) > kCoopLen;
                if (act && !is_long) {
                    int arg = lo;
                    if (hi > lo) {
                        int d = q - lo;
                        uint32_t best = f[lo] + (uint32_t)(d * d);
                        for (int v = lo + 1; v <= hi; ++v) {
                            d = q - v;
                            const uint32_t c = f[v] + (uint32_t)(d * d);
                            if (c < best) { best = c; arg = v; }
                        }
                    }
                    owner[q] = (uint16_t)arg;
                }
                unsigned todo = __ballot_sync(0xffffffffu, is_long);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int qq = __shfl_sync(0xffffffffu, q, src);
                    const int l2 = __shfl_sync(0xffffffffu, lo, src);
                    const int h2 = __shfl_sync(0xffffffffu, hi, src);
                    const int o = coop_owner(f, qq, l2, h2, lane);
                    if (lane == 0) owner[qq] = (uint16_t)o;
                }
            }
        }
        __syncwarp();
    }

    // ---- chained values, in place, 32 pixels at a time (imgproc.h:122-128 incl. its aliasing) ----
    for (int x0 = 0; x0 < n; x0 += 32) {
        const int q = x0 + lane;
        const bool act = q < n;
        const int u = act ? owner[q] : 0;
        const int d = q - u;
        const uint32_t add = (uint32_t)(d * d);
        uint32_t val = 0;
        bool done = !act;
        if (act && (u >= q || u < x0)) {   // right of q (not yet overwritten) or an earlier, finished chunk
            val = f[u] + add;
            done = true;
        }
        // owners inside this chunk and left of q: wait for that lane
        unsigned pending = __ballot_sync(0xffffffffu, !done);
        while (pending) {
            const int src = act ? max(u - x0, 0) : 0;
            const uint32_t sv = __shfl_sync(0xffffffffu, val, src);
            const unsigned dn = __ballot_sync(0xffffffffu, done);
            if (!done && ((dn >> src) & 1u)) {
                val = sv + add;
                done = true;
            }
            pending = __ballot_sync(0xffffffffu, !done);
        }
        __syncwarp();
        if (act) f[q] = val;
        __syncwarp();
    }
    // ---- store (values < 2^24: exact in fp32) ----
    for (int x = lane; x < n; x += 32) out[x] = f[x] >= kBigF ? FLT_MAX : (float)f[x];
}

// K2b (L1): second pass of the L1 transform (core/imgproc.h:137-146,178-184) along x on the u16
// vertical distance; integers are exact, so min-plus order is irrelevant.
__global__ void __launch_bounds__(128) dt_row_l1_kernel(const uint16_t* __restrict__ g, float* __restrict__ planes,
                                                        MapDims dm) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (y >= dm.H) return;
    const size_t base = (size_t)d * dm.plane_elems + (size_t)y * dm.pitch;
    const uint16_t* gin = g + base;
    int* tmp = reinterpret_cast<int*>(planes + base);
    float* out = planes + base;
    const int BIG = 1 << 28;
    int cur = BIG;
    for (int x = 0; x < dm.W; ++x) {
        const int gv = gin[x] == kNoEdge16 ? BIG : (int)gin[x];
        cur = min(gv, cur + 1);
        tmp[x] = cur;
    }
    cur = BIG;
    for (int x = dm.W - 1; x >= 0; --x) {
        cur = min(tmp[x], cur + 1);
        out[x] = cur >= (BIG >> 1) ? FLT_MAX : (float)cur;
    }
}

__global__ void __launch_bounds__(256) sqrt_kernel(float* __restrict__ planes, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) planes[i] = sqrtf(planes[i]);
}

// =============================================================================================
// K3: propagateOrientation (matching/src/featuremaps/dt3cpu.cpp:77-107).  One thread per pixel keeps
// the D-long orientation vector in registers through the forward (ceil(1.5D)) and backward
// (D + floor(1.5D)) circular sweeps; P[c2] = min(P[c2], P[c1] + w_step).  sqrt_first fuses the final
// elementwise sqrt of the L2 transform (core/imgproc.h:191-192) into the load.
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(128) propagate_kernel(float* __restrict__ planes, MapDims dm,
                                                        const __grid_constant__ PropParams pp, int sqrt_first) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dm.plane_elems || (int)(i % dm.pitch) >= dm.W) return;
    float v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = planes[(size_t)d * dm.plane_elems + i];
    if (sqrt_first) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = sqrtf(v[d]);
    }
    constexpr int fwd = (3 * D + 1) / 2;
    constexpr int bwd = D + (3 * D) / 2;
#pragma unroll
    for (int c = 0; c < fwd; ++c) {
        const int c1 = (D + ((c - 1) % D)) % D;
        const int c2 = c % D;
        v[c2] = fminf(v[c2], v[c1] + pp.w[c]);
    }
#pragma unroll
    for (int j = 0; j < bwd; ++j) {
        const int c = D - j;
        const int c1 = (D + ((c + 1) % D)) % D;
        const int c2 = (D + (c % D)) % D;
        v[c2] = fminf(v[c2], v[c1] + pp.w[fwd + j]);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) planes[(size_t)d * dm.plane_elems + i] = v[d];
}

// any depth (orientation vector in local memory)
__global__ void __launch_bounds__(128) propagate_generic_kernel(float* __restrict__ planes, MapDims dm,
                                                                const __grid_constant__ PropParams pp, int sqrt_first) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dm.plane_elems || (int)(i % dm.pitch) >= dm.W) return;
    float v[kMaxDepthDev];
    for (int d = 0; d < dm.D; ++d) {
        const float x = planes[(size_t)d * dm.plane_elems + i];
        v[d] = sqrt_first ? sqrtf(x) : x;
    }
    for (int s = 0; s < pp.n_steps; ++s) {
        const int c1 = pp.c1[s], c2 = pp.c2[s];
        v[c2] = fminf(v[c2], v[c1] + pp.w[s]);
    }
    for (int d = 0; d < dm.D; ++d) planes[(size_t)d * dm.plane_elems + i] = v[d];
}

// =============================================================================================
// K4: lineIntegral (core/imgproc.h:38-84).  A plane's discrete direction (rx, ry) has one unit
// component.  x-major: column i adds column i-1 shifted by dy_i = R(i) - R(i-1), R(j) =
// (long)roundf(j*ry), so pixel (x_i, c + R(i)) continues the chain of pixel (x_{i-1}, c + R(i-1)):
// one thread per chain c carries the strictly sequential fp32 running sum (((a0+a1)+a2)+...).
// y-major is the same with rows/columns swapped (and is the coalesced case for a [H][W] plane).
// =============================================================================================
__global__ void __launch_bounds__(128) integral_kernel(float* __restrict__ planes, MapDims dm,
                                                       const __grid_constant__ IntegralParams ip) {
    const int d = blockIdx.y;
    const int mode = ip.mode[d];
    if (mode == 0) return;
    float* P = planes + (size_t)d * dm.plane_elems;
    const float rx = ip.rx[d], ry = ip.ry[d];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 2) {
        // y-major: rows i = 0..H-1 at y_i = p0y + i*sy; chain c at x_i = c + R(i), R(i) = round(i*rx)
        const int sy = (int)ry;
        const int p0y = ry < 0 ? dm.H - 1 : 0;
        const long long Rend = round_to_ll((float)(dm.H - 1) * rx);
        const long long cmin = Rend > 0 ? -Rend : 0;
        const long long cmax = (Rend < 0 ? -Rend : 0) + dm.W - 1;
        const long long c = cmin + t;
        if (c > cmax) return;
        float acc = 0.f;
        bool have = false;
        for (int i = 0; i < dm.H; ++i) {
            const long long x = c + round_to_ll((float)i * rx);
            if (x < 0 || x >= dm.W) { have = false; continue; }
            float* p = P + (size_t)(p0y + i * sy) * dm.pitch + x;
            const float a = *p;
            if (have) { acc = a + acc; *p = acc; }
            else { acc = a; have = true; }
        }
    } else {
        // x-major: columns i = 0..W-1 at x_i = p0x + i*sx; chain c at y_i = c + R(i), R(i) = round(i*ry)
        const int sx = (int)rx;
        const int p0x = rx < 0 ? dm.W - 1 : 0;
        const long long Rend = round_to_ll((float)(dm.W - 1) * ry);
        const long long cmin = Rend > 0 ? -Rend : 0;
        const long long cmax = (Rend < 0 ? -Rend : 0) + dm.H - 1;
        const long long c = cmin + t;
        if (c > cmax) return;
        float acc = 0.f;
        bool have = false;
        for (int i = 0; i < dm.W; ++i) {
            const long long y = c + round_to_ll((float)i * ry);
            if (y < 0 || y >= dm.H) { have = false; continue; }
            float* p = P + (size_t)y * dm.pitch + (p0x + i * sx);
            const float a = *p;
            if (have) { acc = a + acc; *p = acc; }
            else { acc = a; have = true; }
        }
    }
}

// =============================================================================================
// launchers
// =============================================================================================
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

void launch_raster(const float* d_lines, const int32_t* d_bins, int n_lines, const MapDims& dm, uint32_t* d_mask,
                   cudaStream_t s) {
    if (n_lines <= 0) return;
    raster_kernel<<<cdiv((size_t)n_lines * 32, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(d_lines), d_bins, n_lines,
                                                                   dm, d_mask);
}

void launch_dt_col_exact(const uint32_t* d_mask, const MapDims& dm, uint16_t* d_g, cudaStream_t s) {
    dim3 grid(cdiv(dm.W, 128), dm.D);
    dt_col_exact_kernel<<<grid, 128, 0, s>>>(d_mask, dm, d_g);
}

void launch_mask_to_float(const uint32_t* d_mask, const MapDims& dm, float* d_planes, cudaStream_t s) {
    const size_t total = (size_t)dm.D * dm.plane_elems;
    mask_to_float_kernel<<<cdiv(total, 256), 256, 0, s>>>(d_mask, dm, d_planes);
}

void launch_dt_pass_literal(bool from_g, bool along_rows, const uint16_t* d_g, float* d_planes, const MapDims& dm,
                            void* d_stack, cudaStream_t s) {
    const int n = along_rows ? dm.W : dm.H;
    const int n_lines = along_rows ? dm.H : dm.W;
    const size_t es = along_rows ? 1 : (size_t)dm.pitch;
    const size_t ls = along_rows ? (size_t)dm.pitch : 1;
    dim3 grid(cdiv(n_lines, 128), dm.D);
    if (from_g)
        dt_pass_literal_kernel<true><<<grid, 128, 0, s>>>(d_g, d_planes, dm, (EnvEntry*)d_stack, n, n_lines, es, ls);
    else
        dt_pass_literal_kernel<false><<<grid, 128, 0, s>>>(d_g, d_planes, dm, (EnvEntry*)d_stack, n, n_lines, es, ls);
}

void launch_dt_row_exact(const uint16_t* d_g, float* d_planes, const MapDims& dm, cudaStream_t s) {
    const int warps = 4;
    const size_t smem = (size_t)warps * dm.pitch * (sizeof(uint32_t) + sizeof(uint16_t));
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(dt_row_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const int rows = dm.D * dm.H;
    dt_row_exact_kernel<<<cdiv(rows, warps), warps * 32, smem, s>>>(d_g, d_planes, dm, rows);
}

void launch_dt_row_l1(const uint16_t* d_g, float* d_planes, const MapDims& dm, cudaStream_t s) {
    dim3 grid(cdiv(dm.H, 128), dm.D);
    dt_row_l1_kernel<<<grid, 128, 0, s>>>(d_g, d_planes, dm);
}

void launch_sqrt(float* d_planes, const MapDims& dm, cudaStream_t s) {
    const size_t total = (size_t)dm.D * dm.plane_elems;
    sqrt_kernel<<<cdiv(total, 256), 256, 0, s>>>(d_planes, total);
}

void launch_propagate(float* d_planes, const MapDims& dm, const PropParams& pp, bool sqrt_first, cudaStream_t s) {
    const unsigned grid = cdiv(dm.plane_elems, 128);
    switch (dm.D) {
        case 30: propagate_kernel<30><<<grid, 128, 0, s>>>(d_planes, dm, pp, sqrt_first ? 1 : 0); break;
        case 4: propagate_kernel<4><<<grid, 128, 0, s>>>(d_planes, dm, pp, sqrt_first ? 1 : 0); break;
        default: propagate_generic_kernel<<<grid, 128, 0, s>>>(d_planes, dm, pp, sqrt_first ? 1 : 0); break;
    }
}

void launch_integral(float* d_planes, const MapDims& dm, const IntegralParams& ip, cudaStream_t s) {
    dim3 grid(cdiv((size_t)dm.W + dm.H, 128), dm.D);   // #chains <= W + H
    integral_kernel<<<grid, 128, 0, s>>>(d_planes, dm, ip);
}

}   // namespace fdcm
