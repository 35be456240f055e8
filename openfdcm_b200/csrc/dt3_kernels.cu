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
        // (per call: the attribute is per device and a process may drive several devices)
        if (dm.W % 4 == 0) {
            const size_t smem = (size_t)4 * 2 * kTileRows * kVecPitch * sizeof(float);
            cudaFuncSetAttribute(integral_xmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            integral_xmajor_kernel<<<grid, 128, smem, s>>>(d_planes, dm, ip, d_rtab, rlen);
        } else {   // widths that are no multiple of 4 floats: the 4-byte staging variant
            const size_t smem = (size_t)4 * 2 * kTileRows * kTilePitch * sizeof(float);
            cudaFuncSetAttribute(integral_xmajor_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            integral_xmajor_scalar_kernel<<<grid, 128, smem, s>>>(d_planes, dm, ip, d_rtab, rlen);
        }
    }
}

}   // namespace fdcm
