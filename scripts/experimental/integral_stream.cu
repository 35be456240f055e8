// integral_stream.cu — lineIntegral (reference core/imgproc.h:38-84) with register-staged tile loads (sm_100a).
//
// Same decomposition as integral_tma.cu: a work item is (plane, strip of kSC consecutive chains, range of tiles in which
// the strip meets the image); a CTA walks its strip along the major axis in tiles of kSR steps and every chain is the
// strictly sequential fp32 running sum of the reference (thread = chain).  What differs is how a tile reaches shared
// memory.  The bytes in flight from HBM live in REGISTERS, not in shared-memory stages: all threads of the CTA fetch the
// bounding box of the NEXT tile with 16-byte loads (consecutive threads = consecutive 16-byte pieces of a row) and
// only park them in the single shared-memory tile once the current tile has been summed and written back.  A CTA thus
// needs one tile of shared memory (23 KB) and keeps ~23 KB of loads in flight, five CTAs fit on an SM, and while one
// CTA sums, the loads of all five are outstanding (scripts/micro/rmw_bench.cu: this access pattern streams the planes
// in place at 5.2-6.1 TB/s, the TMA-staged ring of integral_tma.cu saturates at about half of that).
//   y-major: tile = kSR rows x (kSC + kSR + 4) columns; results go from registers to global memory as row segments.
//   x-major: tile = (kSC + kSR) rows x 36 columns (row pitch 36 floats: the column walk of the four 8-lane groups,
//            skewed by 0..3 steps, is bank-conflict free); sums are written back into the tile, then the owned
//            elements leave as row segments (whole 128-byte rows as float4 where every column of the row is owned).
// Box origins are rounded down to 16 bytes; the remainder is added to the shared-memory column.
#include "common.cuh"
#include "kernels.h"

namespace fdcm {

constexpr int kSC = 128;                        // chains per strip = threads per CTA
constexpr int kSR = 32;                         // major-axis steps per tile
constexpr int kSYW = kSC + kSR + 4;             // y-major tile: 32 rows x 164 columns
constexpr int kSYQ = kSYW / 4;                  // 16-byte pieces per row
constexpr int kSXW = 36;                        // x-major tile: 160 rows x 36 columns
constexpr int kSXQ = kSXW / 4;
constexpr int kSXH = kSC + kSR;
constexpr int kSTileFloats = kSXW * kSXH > kSYW * kSR ? kSXW * kSXH : kSYW * kSR;   // 5760 floats
constexpr int kSUnitsY = kSR * kSYQ, kSUnitsX = kSXH * kSXQ;                      // 1312 / 1440 pieces per tile
constexpr int kSUPT = ((kSUnitsX > kSUnitsY ? kSUnitsX : kSUnitsY) + kSC - 1) / kSC;   // pieces per thread: 12
constexpr int kSWarps = kSC / 32;

__global__ void __launch_bounds__(kSC, 5)
integral_stream_kernel(float* __restrict__ planes, MapDims dm, const __grid_constant__ IntegralParams ip,
                       const int32_t* __restrict__ rtab, int rlen, const int4* __restrict__ items, int n_items,
                       int* __restrict__ counter) {
    __shared__ __align__(16) float tile[kSTileFloats];
    __shared__ int s_item;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float4* tile4 = reinterpret_cast<float4*>(tile);
    for (;;) {
        if (tid == 0) s_item = atomicAdd(counter, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= n_items) break;
        const int4 it = items[item];
        const int d = it.x, c0 = it.y, b_lo = it.z, b_hi = it.w;
        const int mode = ip.mode[d];
        // R(i) = (long)roundf(float(i) * r) (imgproc.h:55,72) evaluated in place: the same IEEE multiply and round-half-away
        // as the host table of the plan, and no table load on the path to the tile addresses
        const float rslope = mode == 1 ? ip.ry[d] : ip.rx[d];
        auto R = [&](int i) { return (int)roundf((float)i * rslope); };
        float* P = planes + (size_t)d * dm.plane_elems;
        const int n_major = mode == 1 ? dm.W : dm.H;
        const bool rev = (mode == 1 ? ip.rx[d] : ip.ry[d]) < 0;

        // ---- the bounding box of tile b, 16-byte pieces: consecutive threads take consecutive pieces of a row ----
        float4 st[kSUPT];
        auto fetch = [&](int b) {
            const int i0 = b * kSR, i1 = min(n_major, i0 + kSR);
            const int minor0 = c0 + min(R(i0), R(i1 - 1));                    // R is monotone in i
            if (mode == 2) {
                const int xo = minor0 & ~3;
#pragma unroll
                for (int k = 0; k < kSUPT; ++k) {
                    const int u = tid + k * kSC;
                    const int r = u / kSYQ, q = u - r * kSYQ;
                    const int i = i0 + r, x = xo + 4 * q;
                    const int y = rev ? dm.H - 1 - i : i;
                    st[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u < kSUnitsY && i < i1 && (unsigned)x < (unsigned)dm.pitch)
                        st[k] = *reinterpret_cast<const float4*>(P + (size_t)y * dm.pitch + x);
                }
            } else {
                const int xb = rev ? dm.W - kSR - i0 : i0;
                const int xo = xb & ~3;
#pragma unroll
                for (int k = 0; k < kSUPT; ++k) {
                    const int u = tid + k * kSC;
                    const int r = u / kSXQ, q = u - r * kSXQ;
                    const int y = minor0 + r, x = xo + 4 * q;
                    st[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u < kSUnitsX && (unsigned)y < (unsigned)dm.H && (unsigned)x < (unsigned)dm.pitch)
                        st[k] = *reinterpret_cast<const float4*>(P + (size_t)y * dm.pitch + x);
                }
            }
        };
        auto park = [&]() {
            const int units = mode == 2 ? kSUnitsY : kSUnitsX;
#pragma unroll
            for (int k = 0; k < kSUPT; ++k) {
                const int u = tid + k * kSC;
                if (u < units) tile4[u] = st[k];
            }
        };

        const int c = c0 + tid;
        float acc = 0.f;
        bool have = false;
        if (b_lo < b_hi) fetch(b_lo);
        for (int b = b_lo; b < b_hi; ++b) {
            park();                                                           // (the previous tile was released by the barrier below)
            if (b + 1 < b_hi) fetch(b + 1);                                   // in flight while this tile is summed and stored
            __syncthreads();
            if (mode == 2) {
                // y-major: step i visits row y = i (or H-1-i), the chain is at x = c + R(i)
                const int i0 = b * kSR, nk = min(dm.H - i0, kSR);
                const int Rl = R(min(i0 + lane, dm.H - 1));                   // lane k: shift of step i0 + k
                const int Ra = __shfl_sync(0xffffffffu, Rl, 0), Rb = __shfl_sync(0xffffffffu, Rl, nk - 1);
                const int Rmin = min(Ra, Rb);
                const float* tp = tile + tid - Rmin + ((c0 + Rmin) & 3);      // + k * kSYW + R_k: this chain at step i0 + k
                const long long ystep = rev ? -(long long)dm.pitch : (long long)dm.pitch;
                float* row = P + (long long)(rev ? dm.H - 1 - i0 : i0) * dm.pitch + c;   // + R_k: this chain's pixel of step i0
                // every lane's chain inside the image for the whole tile (x = c + R is monotone along the tile)?
                const bool inside = (unsigned)(c + Ra) < (unsigned)dm.W && (unsigned)(c + Rb) < (unsigned)dm.W;
                const bool fast = __all_sync(0xffffffffu, inside && have);
                if (fast) {
#pragma unroll 8
                    for (int k = 0; k < nk; ++k) {
                        const int Rk = __shfl_sync(0xffffffffu, Rl, k);
                        acc = tp[k * kSYW + Rk] + acc;
                        row[k * ystep + Rk] = acc;
                    }
                } else {
                    for (int k = 0; k < nk; ++k) {
                        const int Rk = __shfl_sync(0xffffffffu, Rl, k);
                        if ((unsigned)(c + Rk) < (unsigned)dm.W) {
                            const float a = tp[k * kSYW + Rk];
                            if (have) { acc = a + acc; row[k * ystep + Rk] = acc; }
                            else { acc = a; have = true; }
                        } else {
                            have = false;
                        }
                    }
                }
            } else {
                // x-major: step i visits column x = i (or W-1-i), the chain is at y = c + R(i)
                const bool fwd = !rev;
                const int grp = lane >> 3;                                    // this lane's lag in the column walk
                const int i0 = b * kSR, ncols = min(dm.W - i0, kSR);
                const int Rl = R(min(i0 + lane, dm.W - 1));
                const int Ra = __shfl_sync(0xffffffffu, Rl, 0), Rb = __shfl_sync(0xffffffffu, Rl, ncols - 1);
                const int Rmin = min(Ra, Rb);
                const int offl = Rl - Rmin;                                   // tile row of chain c0 at step i0 + lane
                const int ybase = c0 + Rmin;
                const int xb = fwd ? i0 : dm.W - kSR - i0;                    // image column of memory column xoff
                const int xoff = xb & 3;                                      // (the box starts at xb & ~3)
                const int mb = xoff + (fwd ? 0 : kSR - 1), ms = fwd ? 1 : -1; // memory column of step j: mb + ms * j
                const bool interior = ncols == kSR && ybase >= 0 && ybase + kSXH <= dm.H;
                const bool fast = __all_sync(0xffffffffu, have) && interior;
                if (fast) {
#pragma unroll 7
                    for (int s = 0; s < kSR + 3; ++s) {
                        const int j = s - grp;
                        const int off = __shfl_sync(0xffffffffu, offl, j & 31);
                        if ((unsigned)j < (unsigned)kSR) {
                            float* e = tile + (tid + off) * kSXW + mb + ms * j;
                            acc = *e + acc;
                            *e = acc;
                        }
                    }
                } else {
                    for (int s = 0; s < kSR + 3; ++s) {
                        const int j = s - grp;
                        const int off = __shfl_sync(0xffffffffu, offl, j & 31);
                        if ((unsigned)j < (unsigned)ncols) {
                            const int r = tid + off;
                            if ((unsigned)(ybase + r) < (unsigned)dm.H) {
                                float* e = tile + r * kSXW + mb + ms * j;
                                if (have) { acc = *e + acc; *e = acc; }
                                else { acc = *e; have = true; }
                            } else {
                                have = false;
                            }
                        }
                    }
                }
                __syncthreads();                                              // every chain of the tile is summed
                // ---- store: tile row r, lane = memory column; the element belongs to chain (r - off of its step) ----
                const int jl = fwd ? lane : kSR - 1 - lane;                   // step of this lane's memory column
                const int off_m = __shfl_sync(0xffffffffu, offl, jl);
                const int x = xb + lane;
                const bool col_ok = jl < ncols;
                if (interior && xoff == 0) {   // (uniform over the CTA: the row distribution below must be the same in every warp)
                    // interior tile: rows [off_max, kSC + off_min) are owned in every column -> whole 128-byte rows as float4
                    // (8 lanes per row, 4 rows per warp instruction); the ragged rows above / below go element-wise
                    const int off_lo = Ra < Rb ? 0 : Ra - Rb, off_hi = Ra < Rb ? Rb - Ra : 0;   // off of the first / last step ...
                    const int full_lo = max(off_lo, off_hi), full_hi = kSC + min(off_lo, off_hi);   // ... off is monotone between them
                    const int sub = lane >> 3, l8 = lane & 7;
                    for (int r = full_lo + warp * 4 + sub; r < full_hi; r += kSWarps * 4) {
                        const float4 v = *reinterpret_cast<const float4*>(tile + r * kSXW + 4 * l8);
                        *reinterpret_cast<float4*>(P + (long long)(ybase + r) * dm.pitch + xb + 4 * l8) = v;
                    }
                    for (int r = warp; r < kSXH; r += kSWarps) {
                        if (r >= full_lo && r < full_hi) { r += ((full_hi - 1 - r) / kSWarps) * kSWarps; continue; }
                        if ((unsigned)(r - off_m) < (unsigned)kSC) P[(long long)(ybase + r) * dm.pitch + x] = tile[r * kSXW + lane];
                    }
                } else {
                    float* gp = P + (long long)(ybase + warp) * dm.pitch + x;
                    const float* tp = tile + warp * kSXW + lane + xoff;
#pragma unroll 4
                    for (int r = warp; r < kSXH; r += kSWarps, gp += (size_t)kSWarps * dm.pitch, tp += kSWarps * kSXW)
                        if (col_ok && (unsigned)(r - off_m) < (unsigned)kSC && (unsigned)(ybase + r) < (unsigned)dm.H) *gp = *tp;
                }
            }
            __syncthreads();                                                  // the tile may be overwritten
        }
    }
}

int integral_stream_chains() { return kSC; }
int integral_stream_steps() { return kSR; }

void launch_integral_stream(float* d_planes, const MapDims& dm, const IntegralParams& ip, const IntegralPlanDev& plan, int n_sms,
                            cudaStream_t s) {
    if (plan.n_items <= 0) return;
    int per_sm = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, integral_stream_kernel, kSC, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    const int grid = plan.n_items < per_sm * n_sms ? plan.n_items : per_sm * n_sms;
    integral_stream_kernel<<<grid, kSC, 0, s>>>(d_planes, dm, ip, plan.rtab, plan.rlen, plan.items4, plan.n_items, plan.counter);
}

}   // namespace fdcm
