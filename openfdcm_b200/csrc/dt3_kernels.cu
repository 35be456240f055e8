// dt3_kernels.cu — hot path 1: DT3 feature-map build kernels (sm_100a).
//
// Pipeline (reference: buildCpuFeaturemap, matching/featuremaps/dt3cpu.h:174-234):
//   K1 raster_kernel            drawLines: clipLines + rasterizeLine per orientation plane  -> 1-bit edge mask
//   K2 dt_col_exact_kernel      first (vertical) pass of distanceTransform on the binary mask -> u16 distance
//      dt_pass_literal_kernel   Felzenszwalb lower-envelope pass, literal incl. the in-place aliasing
//      dt_row_l1_kernel         L1 second pass
//   K3 propagate_kernel         propagateOrientation: 4*D circular min-plus steps, D values in registers
//   K4 integral_tma.cu          lineIntegral: sequential fp32 running sums along each plane's discrete lines
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
// launchers
// =============================================================================================
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

void launch_raster(const float* d_lines, const int32_t* d_bins, int n_lines, const MapDims& dm, uint32_t* d_mask,
                   cudaStream_t s) {
    if (n_lines <= 0) return;
    raster_kernel<<<cdiv((size_t)n_lines * 32, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(d_lines), d_bins, n_lines,
                                                                   dm, d_mask);
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

}   // namespace fdcm
