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
#include <cstdlib>

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

// K2a, tiled: one CTA per (plane, 64 columns).  Phase A transposes the mask into per-column 32-row bit words
// (warp ballots), phase B scans the bands once per column for the nearest edge above / below each band, phase C
// resolves every pixel with two bit scans (clz / ffs) and writes 128-byte row segments of u16 distances.
__global__ void __launch_bounds__(256) dt_col_tiled_kernel(const uint32_t* __restrict__ mask, MapDims dm,
                                                           uint16_t* __restrict__ g, int nbands) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* M = reinterpret_cast<uint32_t*>(smem_raw);                  // [nbands][64] column bit words
    int32_t* up_before = reinterpret_cast<int32_t*>(M + (size_t)nbands * 64);   // [nbands][64] last edge row above the band
    int32_t* dn_after = up_before + (size_t)nbands * 64;                       // [nbands][64] first edge row below the band
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int d = blockIdx.y;
    const int w0 = blockIdx.x * 2;                                        // first mask word (32 columns each)
    const uint32_t* mp = mask + (size_t)d * dm.H * dm.wwords;
    // ---- A: transpose ----
    for (int b = warp; b < nbands; b += nwarps) {
        const int y = b * 32 + lane;
        uint32_t a0 = 0, a1 = 0;
        if (y < dm.H) {
            a0 = mp[(size_t)y * dm.wwords + w0];
            if (w0 + 1 < dm.wwords) a1 = mp[(size_t)y * dm.wwords + w0 + 1];
        }
        uint32_t c0 = 0, c1 = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const uint32_t m0 = __ballot_sync(0xffffffffu, (a0 >> c) & 1u);
            const uint32_t m1 = __ballot_sync(0xffffffffu, (a1 >> c) & 1u);
            if (lane == c) { c0 = m0; c1 = m1; }
        }
        M[(size_t)b * 64 + lane] = c0;
        M[(size_t)b * 64 + 32 + lane] = c1;
    }
    __syncthreads();
    // ---- B: nearest edge row strictly above / below every band, per column ----
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        int last = -1;
        for (int b = 0; b < nbands; ++b) {
            up_before[(size_t)b * 64 + c] = last;
            const uint32_t m = M[(size_t)b * 64 + c];
            if (m) last = b * 32 + 31 - __clz(m);
        }
        int next = -1;
        for (int b = nbands - 1; b >= 0; --b) {
            dn_after[(size_t)b * 64 + c] = next;
            const uint32_t m = M[(size_t)b * 64 + c];
            if (m) next = b * 32 + __ffs(m) - 1;
        }
    }
    __syncthreads();
    // ---- C: per pixel ----
    const int x = blockIdx.x * 64 + lane * 2;                             // this lane's two columns
    if (x >= dm.pitch) return;
    uint16_t* gp = g + (size_t)d * dm.plane_elems + x;
    for (int b = warp; b < nbands; b += nwarps) {
        const uint2 bits = *reinterpret_cast<const uint2*>(M + (size_t)b * 64 + lane * 2);
        const int2 ub = *reinterpret_cast<const int2*>(up_before + (size_t)b * 64 + lane * 2);
        const int2 da = *reinterpret_cast<const int2*>(dn_after + (size_t)b * 64 + lane * 2);
        const int rows = min(32, dm.H - b * 32);
        for (int r = 0; r < rows; ++r) {
            const int y = b * 32 + r;
            const uint32_t le = 0xFFFFFFFFu >> (31 - r), ge = 0xFFFFFFFFu << r;
            uint32_t res = 0;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t m = k ? bits.y : bits.x;
                const int upb = k ? ub.y : ub.x, dna = k ? da.y : da.x;
                const uint32_t above = m & le, below = m & ge;
                const int up = above ? (b * 32 + 31 - __clz(above)) : upb;
                const int dn = below ? (b * 32 + __ffs(below) - 1) : dna;
                unsigned dist = 0xFFFFu;
                if (up >= 0) dist = (unsigned)(y - up);
                if (dn >= 0) dist = min(dist, (unsigned)(dn - y));
                res |= (dist & 0xFFFFu) << (16 * k);
            }
            *reinterpret_cast<uint32_t*>(gp + (size_t)y * dm.pitch) = res;
        }
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
constexpr int kScanLen = 48;              // brackets up to this length are scanned linearly
constexpr int kMaxBlocks = 96;            // 32-column blocks per row (n <= 2897 -> 91)

// leftmost argmin of f[v] + (q-v)^2 over v in [lo, hi] by one lane.  Long brackets are pruned with the
// per-32-column block minima (bmin: min f of the block, bpos: its leftmost position): a block can only hold
// the owner if bmin + dist(q, block)^2 does not exceed the best cost already known.
__device__ __forceinline__ int lane_owner(const uint32_t* f, const uint32_t* bmin, const uint16_t* bpos, int q, int lo, int hi) {
    int d = q - lo;
    uint32_t best = f[lo] + (uint32_t)(d * d);
    int arg = lo;
    if (hi - lo <= kScanLen) {
        for (int v = lo + 1; v <= hi; ++v) {
            d = q - v;
            const uint32_t c = f[v] + (uint32_t)(d * d);
            if (c < best) { best = c; arg = v; }
        }
        return arg;
    }
    // upper bound from valid candidates: lo, hi and the minima of the blocks strictly inside the bracket
    d = q - hi;
    uint32_t U = min(best, f[hi] + (uint32_t)(d * d));
    const int b_lo = lo >> 5, b_hi = hi >> 5;
    for (int b = b_lo + 1; b < b_hi; ++b) {
        d = q - (int)bpos[b];
        U = min(U, bmin[b] + (uint32_t)(d * d));
    }
    // scan, left to right, every block whose lower bound does not exceed the bound (ties must be scanned)
    best = 0xFFFFFFFFu;
    for (int b = b_lo; b <= b_hi; ++b) {
        const int v0 = max(b << 5, lo), v1 = min((b << 5) + 31, hi);
        const int dist = q < v0 ? v0 - q : (q > v1 ? q - v1 : 0);
        if (bmin[b] + (uint32_t)(dist * dist) > U) continue;
        for (int v = v0; v <= v1; ++v) {
            d = q - v;
            const uint32_t c = f[v] + (uint32_t)(d * d);
            if (c < best) { best = c; arg = v; }
        }
        U = min(U, best);
    }
    return arg;
}

// leftmost argmin of f[v] + (q-v)^2 over v in [lo, hi], all 32 lanes cooperating (uniform arguments).
// Blocks whose lower bound bmin + dist(q, block)^2 exceeds the best known cost are skipped.
__device__ int coop_owner(const uint32_t* f, const uint32_t* bmin, const uint16_t* bpos, int q, int lo, int hi, int lane) {
    unsigned long long best = ~0ull;
    if (hi - lo < 96) {
        for (int v = lo + lane; v <= hi; v += 32) {
            const int d = q - v;
            const unsigned long long key = ((unsigned long long)(f[v] + (uint32_t)(d * d)) << 16) | (unsigned)v;
            best = key < best ? key : best;
        }
    } else {
        const int b_lo = lo >> 5, b_hi = hi >> 5;
        int d = q - lo;
        uint32_t U = f[lo] + (uint32_t)(d * d);
        d = q - hi;
        U = min(U, f[hi] + (uint32_t)(d * d));
        for (int b = b_lo + 1 + lane; b < b_hi; b += 32) {   // interior blocks: their minima are valid candidates
            d = q - (int)bpos[b];
            U = min(U, bmin[b] + (uint32_t)(d * d));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) U = min(U, __shfl_xor_sync(0xffffffffu, U, o));
        for (int b0 = b_lo; b0 <= b_hi; b0 += 32) {
            const int b = b0 + lane;
            bool surv = false;
            if (b <= b_hi) {
                const int v0 = max(b << 5, lo), v1 = min((b << 5) + 31, hi);
                const int dist = q < v0 ? v0 - q : (q > v1 ? q - v1 : 0);
                surv = bmin[b] + (uint32_t)(dist * dist) <= U;   // ties must be scanned (leftmost argmin)
            }
            unsigned todo = __ballot_sync(0xffffffffu, surv);
            while (todo) {
                const int v = ((b0 + __ffs(todo) - 1) << 5) + lane;
                todo &= todo - 1;
                if (v >= lo && v <= hi) {
                    d = q - v;
                    const unsigned long long key = ((unsigned long long)(f[v] + (uint32_t)(d * d)) << 16) | (unsigned)v;
                    best = key < best ? key : best;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    return (int)(best & 0xFFFFu);
}

// owners of q and q+1 over [lo, hi] in one cooperative pass (shared loads and block pruning);
// (q+1-v)^2 = (q-v)^2 + 2(q-v) + 1
__device__ void coop_owner_pair(const uint32_t* f, const uint32_t* bmin, const uint16_t* bpos, int q, int lo, int hi, int lane,
                                int& w1, int& w2) {
    unsigned long long best1 = ~0ull, best2 = ~0ull;
    auto eval = [&](int v) {
        const int d = q - v;
        const uint32_t c1 = f[v] + (uint32_t)(d * d);
        const uint32_t c2 = (uint32_t)((int)c1 + 2 * d + 1);
        const unsigned long long k1 = ((unsigned long long)c1 << 16) | (unsigned)v, k2 = ((unsigned long long)c2 << 16) | (unsigned)v;
        best1 = k1 < best1 ? k1 : best1;
        best2 = k2 < best2 ? k2 : best2;
    };
    if (hi - lo < 96) {
        for (int v = lo + lane; v <= hi; v += 32) eval(v);
    } else {
        const int b_lo = lo >> 5, b_hi = hi >> 5;
        int d = q - lo;
        uint32_t c = f[lo] + (uint32_t)(d * d);
        uint32_t U1 = c, U2 = (uint32_t)((int)c + 2 * d + 1);
        d = q - hi;
        c = f[hi] + (uint32_t)(d * d);
        U1 = min(U1, c);
        U2 = min(U2, (uint32_t)((int)c + 2 * d + 1));
        for (int b = b_lo + 1 + lane; b < b_hi; b += 32) {
            d = q - (int)bpos[b];
            c = bmin[b] + (uint32_t)(d * d);
            U1 = min(U1, c);
            U2 = min(U2, (uint32_t)((int)c + 2 * d + 1));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            U1 = min(U1, __shfl_xor_sync(0xffffffffu, U1, o));
            U2 = min(U2, __shfl_xor_sync(0xffffffffu, U2, o));
        }
        for (int b0 = b_lo; b0 <= b_hi; b0 += 32) {
            const int b = b0 + lane;
            bool surv = false;
            if (b <= b_hi) {
                const int v0 = max(b << 5, lo), v1 = min((b << 5) + 31, hi);
                const int dist1 = q < v0 ? v0 - q : (q > v1 ? q - v1 : 0);
                const int dist2 = q + 1 < v0 ? v0 - q - 1 : (q + 1 > v1 ? q + 1 - v1 : 0);
                surv = bmin[b] + (uint32_t)(dist1 * dist1) <= U1 || bmin[b] + (uint32_t)(dist2 * dist2) <= U2;
            }
            unsigned todo = __ballot_sync(0xffffffffu, surv);
            while (todo) {
                const int v = ((b0 + __ffs(todo) - 1) << 5) + lane;
                todo &= todo - 1;
                if (v >= lo && v <= hi) eval(v);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, best1, o), o2 = __shfl_xor_sync(0xffffffffu, best2, o);
        best1 = o1 < best1 ? o1 : best1;
        best2 = o2 < best2 ? o2 : best2;
    }
    w1 = (int)(best1 & 0xFFFFu);
    w2 = (int)(best2 & 0xFFFFu);
}

// leftmost argmin over a short bracket by one lane
__device__ __forceinline__ int scan_owner(const uint32_t* f, int q, int lo, int hi) {
    int d = q - lo;
    uint32_t best = f[lo] + (uint32_t)(d * d);
    int arg = lo;
    for (int v = lo + 1; v <= hi; ++v) {
        d = q - v;
        const uint32_t c = f[v] + (uint32_t)(d * d);
        if (c < best) { best = c; arg = v; }
    }
    return arg;
}

constexpr int kQueueCap = 256;   // intervals in flight per row (power of two)
constexpr uint16_t kUnknown = 0xFFFFu;

// Interval refinement.  pt[q] holds the owner of q at "known" pixels (0xFFFF elsewhere).  An interval (a, b) of known
// pixels with owners oa != ob is split at the crossing x* of the two parabolas oa, ob (the last pixel where oa is
// not worse): because every other parabola minus that two-parabola envelope is convex piecewise linear with its
// minimum at the crossing, a third owner can exist inside (a, b) only if it already wins at x* or x*+1.  So the
// owners of x* and x*+1 (searched over [oa, ob] only, owners are monotone) either certify the boundary or split
// the interval further.  Work is proportional to the number of owner runs, not to the row length.
// Only the columns [win_lo, win_lo + win_w) (multiples of 32; the scene's column range) can hold edges, so only they
// need a slot in the f / out array; pixels outside are never an owner and are written straight to global memory.
__global__ void __launch_bounds__(128) dt_row_exact_kernel(const uint16_t* __restrict__ g, float* __restrict__ planes,
                                                           MapDims dm, int n_rows_total, int win_lo, int win_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = dm.W;
    // per warp: fwin[win_w] u32 | bmin[96] u32 | queue[kQueueCap] u32 | pt[pitch] u16 | bpos[96] u16
    const size_t per_warp = (size_t)win_w * 4 + (size_t)dm.pitch * 2 + kMaxBlocks * 6 + kQueueCap * 4;
    unsigned char* base = smem_raw + (size_t)warp * per_warp;
    uint32_t* fwin = reinterpret_cast<uint32_t*>(base);
    uint32_t* f = fwin - win_lo;                 // absolute column index; only dereferenced inside the window
    uint32_t* bmin = fwin + win_w;
    uint32_t* queue = bmin + kMaxBlocks;
    uint16_t* pt = reinterpret_cast<uint16_t*>(queue + kQueueCap);
    uint16_t* bpos = pt + dm.pitch;
    const int win_hi = win_lo + win_w;           // exclusive
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;   // row of the [D*H][pitch] stack of planes
    if (row >= n_rows_total) return;
    const uint16_t* gin = g + (size_t)row * dm.pitch;
    float* out = planes + (size_t)row * dm.pitch;

    // ---- load g, f = g^2; first / last finite column; clear pt ----
    int cmin = 0x7fffffff, cmax = -1;
    for (int x = lane * 2; x < dm.pitch; x += 64) *reinterpret_cast<uint32_t*>(pt + x) = 0xFFFFFFFFu;
    for (int x = win_lo + lane * 2; x < win_hi; x += 64) {
        const uint32_t two = *reinterpret_cast<const uint32_t*>(gin + x);
        const uint32_t g0 = (x < n) ? (two & 0xFFFFu) : kNoEdge16, g1 = (x + 1 < n) ? (two >> 16) : kNoEdge16;
        f[x] = g0 == kNoEdge16 ? kBigF : g0 * g0;
        f[x + 1] = g1 == kNoEdge16 ? kBigF : g1 * g1;
        if (g0 != kNoEdge16) { cmin = min(cmin, x); cmax = max(cmax, x); }
        if (g1 != kNoEdge16) { cmin = min(cmin, x + 1); cmax = max(cmax, x + 1); }
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
    // ---- per-block minima (lane = block, rotated column order: conflict-free) ----
    for (int b = (win_lo >> 5) + lane; b < (win_hi >> 5); b += 32) {
        uint32_t m = 0xFFFFFFFFu;
        int pos = 0;
        for (int i = 0; i < 32; ++i) {
            const int c = (i + lane) & 31;
            const uint32_t v = f[(b << 5) + c];
            if (v < m || (v == m && c < pos)) { m = v; pos = c; }
        }
        bmin[b] = m;
        bpos[b] = (uint16_t)((b << 5) + pos);
    }
    __syncwarp();

    // ---- owners of the two end pixels, then breadth-first interval refinement ----
    bool overflow = false;
    {
        const int o0 = coop_owner(f, bmin, bpos, 0, cmin, cmax, lane);
        const int o1 = (n > 1) ? coop_owner(f, bmin, bpos, n - 1, o0, cmax, lane) : o0;
        if (lane == 0) {
            pt[0] = (uint16_t)o0;
            pt[n - 1] = (uint16_t)o1;
            queue[0] = (uint32_t)(n - 1) << 16;   // interval (a = 0, b = n-1): a in the low half, b in the high half
        }
    }
    __syncwarp();
    unsigned head = 0, tail = (n > 1) ? 1u : 0u;   // uniform across the warp
    while (head != tail) {
        const unsigned cnt = min(32u, tail - head);
        const bool act = (unsigned)lane < cnt;
        int a = 0, b = 0, oa = 0, ob = 0;
        if (act) {
            const uint32_t e = queue[(head + lane) & (kQueueCap - 1)];
            a = (int)(e & 0xFFFFu);
            b = (int)(e >> 16);
            oa = pt[a];
            ob = pt[b];
        }
        head += cnt;
        // an interval needs work only if its end owners differ and it has interior pixels
        const bool work = act && oa != ob && b > a + 1;
        int x = 0, w1 = 0, w2 = 0;
        bool is_long = false;
        if (work) {
            // last pixel where parabola oa is not worse than ob: floor((f_b - f_a + ob^2 - oa^2) / (2 (ob - oa)))
            const int num = (int)f[ob] - (int)f[oa] + ob * ob - oa * oa;
            const int den = 2 * (ob - oa);
            int xs = num / den;
            if (num % den != 0 && num < 0) --xs;
            x = min(max(xs, a), b - 1);
            is_long = (ob - oa) > kScanLen;
            if (!is_long) {
                w1 = (x == a) ? oa : scan_owner(f, x, oa, ob);
                w2 = (x + 1 == b) ? ob : scan_owner(f, x + 1, w1, ob);
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, is_long);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int xq = __shfl_sync(0xffffffffu, x, src);
            const int la = __shfl_sync(0xffffffffu, a, src), lb = __shfl_sync(0xffffffffu, b, src);
            const int l2 = __shfl_sync(0xffffffffu, oa, src), h2 = __shfl_sync(0xffffffffu, ob, src);
            int r1, r2;
            if (xq == la) {
                r1 = l2;
                r2 = (xq + 1 == lb) ? h2 : coop_owner(f, bmin, bpos, xq + 1, l2, h2, lane);
            } else if (xq + 1 == lb) {
                r2 = h2;
                r1 = coop_owner(f, bmin, bpos, xq, l2, h2, lane);
            } else {
                coop_owner_pair(f, bmin, bpos, xq, l2, h2, lane, r1, r2);
            }
            if (lane == src) { w1 = r1; w2 = r2; }
        }
        // publish the two new known pixels and enqueue the children that still straddle a boundary
        bool c1 = false, c2 = false;
        if (work) {
            pt[x] = (uint16_t)w1;
            pt[x + 1] = (uint16_t)w2;
            c1 = (w1 != oa) && (x > a + 1);
            c2 = (w2 != ob) && (b > x + 2);
        }
        const unsigned m1 = __ballot_sync(0xffffffffu, c1), m2 = __ballot_sync(0xffffffffu, c2);
        const unsigned n1 = __popc(m1), n2 = __popc(m2);
        if (tail - head + n1 + n2 > (unsigned)kQueueCap) { overflow = true; break; }
        const unsigned lt = (1u << lane) - 1u;
        if (c1) queue[(tail + __popc(m1 & lt)) & (kQueueCap - 1)] = (uint32_t)a | ((uint32_t)x << 16);
        if (c2) queue[(tail + n1 + __popc(m2 & lt)) & (kQueueCap - 1)] = (uint32_t)(x + 1) | ((uint32_t)b << 16);
        tail += n1 + n2;
        __syncwarp();
    }
    if (overflow) {
        // more than kQueueCap open intervals (very dense rows): plain divide and conquer over every pixel
        __syncwarp();
        for (int q = 31 + 32 * lane; q < n; q += 1024) pt[q] = (uint16_t)lane_owner(f, bmin, bpos, q, cmin, cmax);
        __syncwarp();
        for (int s = 16; s >= 1; s >>= 1) {
            for (int q = s - 1 + 2 * s * lane; q < n; q += 64 * s) {
                const int lo = (q - s >= 0) ? pt[q - s] : cmin;
                const int hi = (q + s < n) ? pt[q + s] : cmax;
                pt[q] = (uint16_t)((lo == hi) ? lo : lane_owner(f, bmin, bpos, q, lo, hi));
            }
            __syncwarp();
        }
    }
    __syncwarp();

    // ---- chained values, in place, 32 pixels at a time (imgproc.h:122-128 incl. its aliasing) ----
    int carry = 0;   // owner of the last known pixel of the previous chunks (pixel 0 is always known)
    for (int x0 = 0; x0 < n; x0 += 32) {
        const int q = x0 + lane;
        const bool act = q < n;
        // owner(q) = owner of the nearest known pixel at or left of q
        const unsigned pv = act ? pt[q] : kUnknown;
        const unsigned known = __ballot_sync(0xffffffffu, pv != kUnknown);
        if (known == 0) {
            // whole chunk inside one run (owner = carry): one broadcast read, no dependencies.  If the owner pixel is
            // in this chunk it owns itself, so its slot is rewritten with the same value (f[carry] + 0)
            if (act) {
                const int dd = q - carry;
                const uint32_t vv = f[carry] + (uint32_t)(dd * dd);
                out[q] = (float)vv;
                if (q >= win_lo && q < win_hi) f[q] = vv;
            }
            __syncwarp();
            continue;
        }
        const unsigned le = known & (0xFFFFFFFFu >> (31 - lane));
        const int srcl = le ? 31 - __clz(le) : 0;
        const int ul = __shfl_sync(0xffffffffu, (int)pv, srcl);
        const int u = le ? ul : carry;
        if (known) carry = __shfl_sync(0xffffffffu, (int)pv, 31 - __clz(known));
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
        if (act) {
            out[q] = (float)val;                              // < 2^24: exact in fp32
            if (q >= win_lo && q < win_hi) f[q] = val;        // only window columns can be referenced again
        }
        __syncwarp();
    }
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

// K2b (L1), warp per row: out[x] = min(x + min_{v<=x}(g[v]-v), -x + min_{v>=x}(g[v]+v)) with warp prefix / suffix
// min scans (integers: exact); the forward result is parked in shared memory between the two sweeps.
__global__ void __launch_bounds__(128) dt_row_l1_warp_kernel(const uint16_t* __restrict__ g, float* __restrict__ planes,
                                                             MapDims dm, int n_rows_total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* fwd = reinterpret_cast<int*>(smem_raw) + (size_t)warp * dm.pitch;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= n_rows_total) return;
    const int n = dm.W;
    const uint16_t* gin = g + (size_t)row * dm.pitch;
    float* out = planes + (size_t)row * dm.pitch;
    const int BIG = 1 << 28;
    int carry = BIG;   // min over previous chunks of g[v] - v
    for (int x0 = 0; x0 < n; x0 += 32) {
        const int x = x0 + lane;
        int a = BIG;
        if (x < n) {
            const int gv = gin[x];
            a = gv == kNoEdge16 ? BIG : gv - x;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, a, o);
            if (lane >= o) a = min(a, t);
        }
        a = min(a, carry);
        if (x < n) fwd[x] = a;
        carry = __shfl_sync(0xffffffffu, a, 31);
    }
    __syncwarp();
    carry = BIG;       // min over later chunks of g[v] + v
    for (int x0 = ((n - 1) >> 5) << 5; x0 >= 0; x0 -= 32) {
        const int x = x0 + lane;
        int b = BIG;
        if (x < n) {
            const int gv = gin[x];
            b = gv == kNoEdge16 ? BIG : gv + x;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, b, o);
            if (lane + o < 32) b = min(b, t);
        }
        b = min(b, carry);
        if (x < n) {
            const int v = min(fwd[x] + x, b - x);
            out[x] = v >= (BIG >> 1) ? FLT_MAX : (float)v;
        }
        carry = __shfl_sync(0xffffffffu, b, 0);
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
// R(i) = (long)roundf(float(i) * r) per plane (r = ry for x-major, rx for y-major planes): the cumulative
// minor-axis shift of a chain after i major-axis steps (sum of the reference's per-step deltas, imgproc.h:55,72)
__global__ void integral_shift_table_kernel(int32_t* __restrict__ rtab, int len, const __grid_constant__ IntegralParams ip) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (i >= len) return;
    const float r = ip.mode[d] == 1 ? ip.ry[d] : ip.rx[d];
    rtab[(size_t)d * len + i] = (int32_t)round_to_ll((float)i * r);
}

__device__ __forceinline__ void cp_async_f32(uint32_t smem_addr, const float* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// y-major planes: thread per chain, the row access of a warp is one coalesced 128-byte segment; kYU loads in
// flight per thread hide the HBM latency (the running sum itself is the only serial dependency).
constexpr int kYU = 32;
constexpr int kYPrefetch = 96;          // rows of L2 prefetch distance

__global__ void __launch_bounds__(128) integral_ymajor_kernel(float* __restrict__ planes, MapDims dm,
                                                              const __grid_constant__ IntegralParams ip,
                                                              const int32_t* __restrict__ rtab, int rlen) {
    const int d = blockIdx.y;
    if (ip.mode[d] != 2) return;
    const int32_t* R = rtab + (size_t)d * rlen;
    const float ry = ip.ry[d];
    const int Rend = R[dm.H - 1];
    const int cmin = Rend > 0 ? -Rend : 0;
    const int cmax = (Rend < 0 ? -Rend : 0) + dm.W - 1;
    int c = cmin + blockIdx.x * blockDim.x + threadIdx.x;
    if (cmin + (int)(blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) > cmax) return;   // whole warp past the last chain
    if (c > cmax) c = 0x20000000;                                                         // idle lane: every x is out of range
    // rows are visited at y = p0y + i*sy: fold the direction into a row pointer and a signed pitch
    const long long rstep = ry < 0 ? -(long long)dm.pitch : (long long)dm.pitch;
    float* row = planes + (size_t)d * dm.plane_elems + (ry < 0 ? (size_t)(dm.H - 1) * dm.pitch : 0);
    float acc = 0.f;
    bool have = false;
    const int lane = threadIdx.x & 31;
    const int cw = cmin + (int)(blockIdx.x * blockDim.x + (threadIdx.x & ~31u));   // first chain of the warp
    for (int i0 = 0; i0 < dm.H; i0 += kYU) {
        float a[kYU];
        const int32_t rl = (i0 + lane < dm.H) ? __ldg(R + i0 + lane) : 0x40000000;   // lane k: shift of row i0+k
        {   // pull the rows kYPrefetch ahead into L2: lane k takes row i0 + kYPrefetch + k (the warp's 128-byte segment
            // of that row starts at chain cw; it may straddle two lines)
            const int ip = i0 + kYPrefetch + lane;
            if (ip < dm.H) {
                const int xp = cw + __ldg(R + ip);
                const float* rowp = row + (long long)(kYPrefetch + lane) * rstep;
                if ((unsigned)xp < (unsigned)dm.W) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + xp));
                if ((unsigned)(xp + 31) < (unsigned)dm.W) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + xp + 31));
            }
        }
#pragma unroll
        for (int k = 0; k < kYU; ++k) {
            const int x = c + __shfl_sync(0xffffffffu, rl, k);
            a[k] = ((unsigned)x < (unsigned)dm.W) ? row[(long long)k * rstep + x] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kYU; ++k) {
            const int x = c + __shfl_sync(0xffffffffu, rl, k);
            if ((unsigned)x < (unsigned)dm.W) {
                if (have) { acc = a[k] + acc; row[(long long)k * rstep + x] = acc; }
                else { acc = a[k]; have = true; }
            } else {
                have = false;
            }
        }
        row += (long long)kYU * rstep;
    }
}

// x-major planes: a warp owns 32 consecutive chains and walks the columns in blocks of 32.  Each block is
// staged through a shared-memory tile (<= 64 rows x 32 columns, row segments loaded / stored as coalesced
// 128-byte pieces), lane j then runs chain j sequentially across the 32 columns of the tile.  Tiles are double
// buffered with cp.async: the next tile's rows are in flight while the current one is summed and written back.
constexpr int kTileRows = 64, kTilePitch = 33;

struct XTile {              // geometry of one 32-column block for one warp
    int ybase, r_lo, r_hi, off, ncols;
};

__global__ void __launch_bounds__(128) integral_xmajor_scalar_kernel(float* __restrict__ planes, MapDims dm,
                                                              const __grid_constant__ IntegralParams ip,
                                                              const int32_t* __restrict__ rtab, int rlen) {
    extern __shared__ __align__(16) float tiles_all[];       // [4 warps][2 buffers][kTileRows * kTilePitch]
    const int d = blockIdx.y;
    if (ip.mode[d] != 1) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile0 = tiles_all + (size_t)warp * 2 * kTileRows * kTilePitch;
    const uint32_t tile0_s = (uint32_t)__cvta_generic_to_shared(tile0);
    float* P = planes + (size_t)d * dm.plane_elems;
    const int32_t* R = rtab + (size_t)d * rlen;
    const float rx = ip.rx[d];
    const int sx = rx < 0 ? -1 : 1;
    const int p0x = rx < 0 ? dm.W - 1 : 0;
    const int Rend = R[dm.W - 1];
    const int cmin = Rend > 0 ? -Rend : 0;
    const int cmax = (Rend < 0 ? -Rend : 0) + dm.H - 1;
    const int c0 = cmin + (blockIdx.x * 4 + warp) * 32;                    // first chain of this warp
    if (c0 > cmax) return;
    auto geometry = [&](int i0) {
        XTile g;
        const int i = i0 + lane;                                           // this lane's column step (load/store role)
        const bool col_ok = i < dm.W;
        const int Rl = R[col_ok ? i : dm.W - 1];
        const int Ra = __shfl_sync(0xffffffffu, Rl, 0);
        const int Rb = __shfl_sync(0xffffffffu, Rl, min(31, dm.W - 1 - i0));
        const int Rmin = min(Ra, Rb);
        g.ybase = c0 + Rmin;
        const int nrows = 32 + abs(Ra - Rb);
        g.r_lo = max(0, -g.ybase);                                         // rows of the tile inside the image
        g.r_hi = min(nrows, dm.H - g.ybase);
        g.off = col_ok ? Rl - Rmin : 0x40000000;                           // tile row of chain c0 in this lane's column
        g.ncols = min(32, dm.W - i0);
        return g;
    };
    // row r of the tile, lane = column; the element belongs to chain (r - off)
    auto load = [&](int i0, int buf) {
        if (i0 < dm.W) {
            const XTile g = geometry(i0);
            const float* gp = P + (long long)(g.ybase + g.r_lo) * dm.pitch + (p0x + (i0 + lane) * sx);
            uint32_t sp = tile0_s + (uint32_t)((buf * kTileRows + g.r_lo) * kTilePitch + lane) * 4u;
            for (int r = g.r_lo; r < g.r_hi; ++r, gp += dm.pitch, sp += kTilePitch * 4u)
                if ((unsigned)(r - g.off) < 32u) cp_async_f32(sp, gp);
        }
        cp_async_commit();
    };
    float acc = 0.f;
    bool have = false;
    load(0, 0);
    int buf = 0;
    for (int i0 = 0; i0 < dm.W; i0 += 32, buf ^= 1) {
        load(i0 + 32, buf ^ 1);
        cp_async_wait<1>();
        __syncwarp();
        float* tile = tile0 + (size_t)buf * kTileRows * kTilePitch;
        const XTile g = geometry(i0);
        // ---- sequential sums: lane = chain; the 32 tile values of the chain go through registers ----
        float v[32];
        int rr[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int r = lane + __shfl_sync(0xffffffffu, g.off, j);       // tile row of this lane's chain in column j
            rr[j] = (j < g.ncols && r >= g.r_lo && r < g.r_hi) ? r * kTilePitch + j : -1;
            v[j] = rr[j] >= 0 ? tile[rr[j]] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (rr[j] >= 0) {
                if (have) { acc = v[j] + acc; tile[rr[j]] = acc; }
                else { acc = v[j]; have = true; }
            } else if (j < g.ncols) {
                have = false;
            }
        }
        __syncwarp();
        // ---- store ----
        float* gp = P + (long long)(g.ybase + g.r_lo) * dm.pitch + (p0x + (i0 + lane) * sx);
        for (int r = g.r_lo; r < g.r_hi; ++r, gp += dm.pitch)
            if ((unsigned)(r - g.off) < 32u) *gp = tile[r * kTilePitch + lane];
        __syncwarp();
    }
}

// x-major planes, vector path (W % 4 == 0): the same tiles, but staged with 16-byte cp.async (LDGSTS.128 moves 512
// bytes per warp instruction; the 4-byte form is limited to about one element per cycle per SM) as full row segments
// in memory order (pitch kVecPitch floats, 16-byte aligned rows).  A column walk over 16-byte staged rows can only
// reach the 8 banks congruent to the column (mod 4), so the four 8-lane groups of the warp run 0..3 columns behind
// each other: bank = 4*(lane%8) + 4*shift + column - lane/8 is then distinct for all 32 lanes.
constexpr int kVecPitch = 36;
constexpr int kSkew = 3;

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const float* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

__global__ void __launch_bounds__(128) integral_xmajor_kernel(float* __restrict__ planes, MapDims dm,
                                                              const __grid_constant__ IntegralParams ip,
                                                              const int32_t* __restrict__ rtab, int rlen) {
    extern __shared__ __align__(16) float tiles_all[];       // [4 warps][2 buffers][kTileRows * kVecPitch]
    const int d = blockIdx.y;
    if (ip.mode[d] != 1) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile0 = tiles_all + (size_t)warp * 2 * kTileRows * kVecPitch;
    const uint32_t tile0_s = (uint32_t)__cvta_generic_to_shared(tile0);
    float* P = planes + (size_t)d * dm.plane_elems;
    const int32_t* R = rtab + (size_t)d * rlen;
    const bool fwd = !(ip.rx[d] < 0);                                      // column step i is x = i (fwd) or x = W-1-i
    const int Rend = R[dm.W - 1];
    const int cmin = Rend > 0 ? -Rend : 0;
    const int cmax = (Rend < 0 ? -Rend : 0) + dm.H - 1;
    const int c0 = cmin + (blockIdx.x * 4 + warp) * 32;                    // first chain of this warp
    if (c0 > cmax) return;
    const int grp = lane >> 3;                                             // this lane's lag in the column walk
    auto geometry = [&](int i0) {
        XTile g;
        const int i = i0 + lane;                                           // this lane's column step (shift-table role)
        const bool col_ok = i < dm.W;
        const int Rl = R[col_ok ? i : dm.W - 1];
        const int Ra = __shfl_sync(0xffffffffu, Rl, 0);
        const int Rb = __shfl_sync(0xffffffffu, Rl, min(31, dm.W - 1 - i0));
        const int Rmin = min(Ra, Rb);
        g.ybase = c0 + Rmin;
        const int nrows = 32 + abs(Ra - Rb);
        g.r_lo = max(0, -g.ybase);                                         // rows of the tile inside the image
        g.r_hi = min(nrows, dm.H - g.ybase);
        g.off = col_ok ? Rl - Rmin : 0x40000000;                           // tile row of chain c0 at this lane's column step
        g.ncols = min(32, dm.W - i0);
        return g;
    };
    // memory column m of the tile <-> x = xbase + m; column step j sits at m = j (fwd) or 31 - j
    auto load = [&](int i0, int buf) {
        if (i0 < dm.W) {
            const XTile g = geometry(i0);
            const int xbase = fwd ? i0 : dm.W - 32 - i0;
            const int x = xbase + 4 * (lane & 7);
            const bool chunk_ok = x >= 0 && x + 3 < dm.W;
            const int r0 = g.r_lo + (lane >> 3);
            const float* gp = P + (long long)(g.ybase + r0) * dm.pitch + x;
            uint32_t sp = tile0_s + (uint32_t)((buf * kTileRows + r0) * kVecPitch + 4 * (lane & 7)) * 4u;
            for (int r = r0; r < g.r_hi; r += 4, gp += 4 * (long long)dm.pitch, sp += 4u * kVecPitch * 4u)
                if (chunk_ok) cp_async_16(sp, gp);
        }
        cp_async_commit();
    };
    float acc = 0.f;
    bool have = false;
    load(0, 0);
    int buf = 0;
    for (int i0 = 0; i0 < dm.W; i0 += 32, buf ^= 1) {
        load(i0 + 32, buf ^ 1);
        cp_async_wait<1>();
        __syncwarp();
        float* tile = tile0 + (size_t)buf * kTileRows * kVecPitch;
        const XTile g = geometry(i0);
        // ---- sequential sums: lane = chain, step s handles column step j = s - grp ----
        const int mb = fwd ? 0 : 31, ms = fwd ? 1 : -1;                   // memory column of column step j: mb + ms * j
        const bool interior = g.ncols == 32 && g.r_lo == 0 && g.ybase + 64 <= dm.H;   // every chain element of the tile exists
        if (interior && __all_sync(0xffffffffu, have)) {
            float v[32 + kSkew];
            int rr[32 + kSkew];
#pragma unroll
            for (int s = 0; s < 32 + kSkew; ++s) {
                const int j = s - grp;
                if (s >= kSkew && s < 32) {                                // all four groups are inside the tile
                    rr[s] = (lane + __shfl_sync(0xffffffffu, g.off, j)) * kVecPitch + mb + ms * j;
                    v[s] = tile[rr[s]];
                } else {
                    rr[s] = (lane + __shfl_sync(0xffffffffu, g.off, j & 31)) * kVecPitch + mb + ms * j;
                    v[s] = (unsigned)j < 32u ? tile[rr[s]] : 0.f;
                }
            }
#pragma unroll
            for (int s = 0; s < 32 + kSkew; ++s) {
                if (s >= kSkew && s < 32) {
                    acc = v[s] + acc;
                    tile[rr[s]] = acc;
                } else if ((unsigned)(s - grp) < 32u) {
                    acc = v[s] + acc;
                    tile[rr[s]] = acc;
                }
            }
        } else {
            float v[32 + kSkew];
            int rr[32 + kSkew];
#pragma unroll
            for (int s = 0; s < 32 + kSkew; ++s) {
                const int j = s - grp;
                const int r = lane + __shfl_sync(0xffffffffu, g.off, j & 31);  // tile row of this lane's chain at column step j
                const bool ok = (unsigned)j < (unsigned)g.ncols && r >= g.r_lo && r < g.r_hi;
                rr[s] = ok ? r * kVecPitch + mb + ms * j : ((unsigned)j < (unsigned)g.ncols ? -1 : -2);
                v[s] = ok ? tile[rr[s]] : 0.f;
            }
#pragma unroll
            for (int s = 0; s < 32 + kSkew; ++s) {
                if (rr[s] >= 0) {
                    if (have) { acc = v[s] + acc; tile[rr[s]] = acc; }
                    else { acc = v[s]; have = true; }
                } else if (rr[s] == -1) {                                  // a column step of the image outside the plane
                    have = false;
                }
            }
        }
        __syncwarp();
        // ---- store: row r of the tile, lane = memory column; the element belongs to chain (r - off) ----
        const int j_of_lane = fwd ? lane : 31 - lane;
        const int off_m = __shfl_sync(0xffffffffu, g.off, j_of_lane);
        const int xbase = fwd ? i0 : dm.W - 32 - i0;
        float* gp = P + (long long)(g.ybase + g.r_lo) * dm.pitch + (xbase + lane);
        for (int r = g.r_lo; r < g.r_hi; ++r, gp += dm.pitch)
            if ((unsigned)(r - off_m) < 32u) *gp = tile[r * kVecPitch + lane];
        __syncwarp();
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
    const int nbands = (dm.H + 31) / 32;
    const size_t smem = (size_t)nbands * 64 * 12;
    if (smem <= 200 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(dt_col_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set = true;
        }
        dim3 grid((dm.wwords + 1) / 2, dm.D);
        dt_col_tiled_kernel<<<grid, 256, smem, s>>>(d_mask, dm, d_g, nbands);
        return;
    }
    dim3 grid(cdiv(dm.W, 128), dm.D);   // very tall maps: one thread per column, two sweeps
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

void launch_dt_row_exact(const uint16_t* d_g, float* d_planes, const MapDims& dm, int col_lo, int col_hi, cudaStream_t s) {
    // column window that can hold edge pixels, widened to multiples of 32
    int win_lo = (col_lo < 0 ? 0 : col_lo) & ~31;
    int win_hi = ((col_hi >= dm.W ? dm.W - 1 : col_hi) + 32) & ~31;
    if (win_hi > dm.pitch) win_hi = dm.pitch;
    if (win_hi <= win_lo) { win_lo = 0; win_hi = dm.pitch; }
    const int win_w = win_hi - win_lo;
    const size_t per_warp = (size_t)win_w * 4 + (size_t)dm.pitch * 2 + 96 * 6 + 256 * 4;
    // as many rows in flight per SM as shared memory allows (227 KB, 1 KB reserved per CTA)
    int warps = 4;
    int best_rows = 0;
    for (int w = 1; w <= 4; ++w) {
        const int ctas = (int)((227 * 1024) / (per_warp * w + 1024));
        if (ctas * w > best_rows) { best_rows = ctas * w; warps = w; }
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(dt_row_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    const int rows = dm.D * dm.H;
    dt_row_exact_kernel<<<cdiv(rows, warps), warps * 32, per_warp * warps, s>>>(d_g, d_planes, dm, rows, win_lo, win_w);
}

void launch_dt_row_l1(const uint16_t* d_g, float* d_planes, const MapDims& dm, cudaStream_t s) {
    const size_t per_warp = (size_t)dm.pitch * 4;
    if (per_warp * 4 <= 200 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(dt_row_l1_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_set = true;
        }
        const int rows = dm.D * dm.H;
        dt_row_l1_warp_kernel<<<cdiv(rows, 4), 128, per_warp * 4, s>>>(d_g, d_planes, dm, rows);
        return;
    }
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

void launch_integral(float* d_planes, const MapDims& dm, const IntegralParams& ip, int32_t* d_rtab, cudaStream_t s) {
    const int rlen = dm.W > dm.H ? dm.W : dm.H;
    {
        dim3 tgrid(cdiv(rlen, 256), dm.D);
        integral_shift_table_kernel<<<tgrid, 256, 0, s>>>(d_rtab, rlen, ip);
    }
    dim3 grid(cdiv((size_t)dm.W + dm.H, 128), dm.D);   // #chains <= W + H
    bool any_x = false, any_y = false;
    for (int d = 0; d < dm.D; ++d) {
        any_x |= ip.mode[d] == 1;
        any_y |= ip.mode[d] == 2;
    }
    if (any_y) integral_ymajor_kernel<<<grid, 128, 0, s>>>(d_planes, dm, ip, d_rtab, rlen);
    if (any_x) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(integral_xmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(4 * 2 * kTileRows * kVecPitch * sizeof(float)));
            cudaFuncSetAttribute(integral_xmajor_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(4 * 2 * kTileRows * kTilePitch * sizeof(float)));
            attr_set = true;
        }
        static const bool force_scalar = [] { const char* e = getenv("FDCM_INTEGRAL_SCALAR"); return e && e[0] == '1'; }();
        if (dm.W % 4 == 0 && !force_scalar)
            integral_xmajor_kernel<<<grid, 128, (size_t)4 * 2 * kTileRows * kVecPitch * sizeof(float), s>>>(d_planes, dm, ip, d_rtab, rlen);
        else
            integral_xmajor_scalar_kernel<<<grid, 128, (size_t)4 * 2 * kTileRows * kTilePitch * sizeof(float), s>>>(d_planes, dm, ip, d_rtab,
                                                                                                                 rlen);
    }
}

}   // namespace fdcm
