// integral_tma.cu — lineIntegral (reference core/imgproc.h:38-84) as one persistent, TMA-fed kernel (sm_100a).
//
// A plane's discrete direction (rx, ry) has one unit component.  x-major: column step i adds column i-1 shifted by
// dy_i = R(i) - R(i-1), R(j) = (long)roundf(j * ry), so pixel (x_i, c + R(i)) continues the chain of pixel
// (x_{i-1}, c + R(i-1)); y-major is the same with rows and columns swapped.  Every chain is the strictly sequential
// fp32 running sum (((a0 + a1) + a2) + ...) of the reference: parallelism is across chains only.
//
// Work item = (plane, strip of kICW consecutive chains).  A CTA walks its strip along the major axis in tiles of kIRB
// steps.  One elected producer thread stages the tiles with TMA box loads (cp.async.bulk.tensor -> UTMALDG) into a
// kIStages-deep shared-memory ring guarded by full / empty mbarriers, so several tiles per CTA are in flight from HBM
// while four consumer warps (thread = chain) run the sums out of shared memory.  Out-of-image parts of a box are
// zero-filled by the TMA unit, which removes every load-side bounds check.
//   y-major: box = kIRB rows x (kICW + kIRB + 4) columns; the lanes of a warp read / write consecutive x: results go
//            straight from registers to global memory as 128-byte row segments.
//   (the innermost box coordinate of a tensor load must be a multiple of 16 bytes: box origins are rounded down to 4
//   floats and the remainder is added to the shared-memory column)
//   x-major: box = (kICW + kIRB) rows x 36 columns (32 steps; the row pitch of 36 floats makes the column walk of the
//            four 8-lane groups, skewed by 0..3 steps, bank-conflict free); sums are written back into the tile, then
//            the owned elements leave as row segments.
// Items are handed out through an atomic counter in decreasing order of work (persistent CTAs, as many per SM as fit).
#include "common.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace fdcm {

constexpr int kICW = 128;                       // chains per strip = consumer threads
constexpr int kIRB = 32;                        // major-axis steps per tile
constexpr int kIYBoxW = kICW + kIRB + 4;        // y-major box: 164 columns x 32 rows (the box origin is rounded down to 16 bytes)
constexpr int kIXBoxW = 36;                     // x-major box: 36 columns x 160 rows
constexpr int kIXBoxH = kICW + kIRB;
constexpr int kIStages = 2;
constexpr int kIStageFloats = kIXBoxW * kIXBoxH > kIYBoxW * kIRB ? kIXBoxW * kIXBoxH : kIYBoxW * kIRB;   // 5760 floats
constexpr int kIConsumerWarps = kICW / 32;
constexpr int kIThreads = kICW + 32;            // + the producer warp
static_assert((kIStageFloats * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");

__global__ void __launch_bounds__(kIThreads, 4)
integral_tma_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x,
                    float* __restrict__ planes, MapDims dm, const __grid_constant__ IntegralParams ip,
                    const int32_t* __restrict__ rtab, int rlen, const int2* __restrict__ items, int n_items,
                    int* __restrict__ counter) {
    extern __shared__ __align__(128) float stages[];          // [kIStages][kIStageFloats]
    __shared__ __align__(8) uint64_t bars[2 * kIStages];      // full[0..S), empty[S..2S)
    __shared__ int s_item;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kIStages);
    const uint32_t stage0 = smem_u32(stages);
    if (tid == 0) {
        for (int s = 0; s < kIStages; ++s) {
            mbar_init(full0 + 8u * s, 1);                     // the producer's arrive.expect_tx
            mbar_init(empty0 + 8u * s, kIConsumerWarps);      // one arrival per consumer warp
        }
        mbar_init_fence();
        prefetch_tensormap(&map_y);
        prefetch_tensormap(&map_x);
    }
    __syncthreads();
    uint32_t it = 0;                                          // tiles handled so far (same count in every thread)
    for (;;) {
        if (tid == 0) s_item = atomicAdd(counter, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= n_items) break;
        const int d = items[item].x, c0 = items[item].y;
        const int mode = ip.mode[d];
        const int32_t* R = rtab + (size_t)d * rlen;
        float* P = planes + (size_t)d * dm.plane_elems;
        const int n_major = mode == 1 ? dm.W : dm.H;
        const int nblk = (n_major + kIRB - 1) / kIRB;

        if (warp == kIConsumerWarps) {
            // ---------------- producer: one thread issues every box load of the item ----------------
            if (lane == 0) {
                const bool rev = (mode == 1 ? ip.rx[d] : ip.ry[d]) < 0;
                // box origin of tile b: (innermost coordinate rounded down to 16 bytes, second coordinate)
                auto origin = [&](int b, int& cx, int& cy) {
                    const int i0 = b * kIRB, i1 = min(n_major, i0 + kIRB);
                    const int minor0 = c0 + min(R[i0], R[i1 - 1]);            // R is monotone in i
                    const int major0 = rev ? n_major - kIRB - i0 : i0;
                    if (mode == 1) { cx = major0 & ~3; cy = minor0; }
                    else { cx = minor0 & ~3; cy = major0; }
                };
                const CUtensorMap* map = mode == 1 ? &map_x : &map_y;
                const uint32_t bytes = mode == 1 ? kIXBoxW * kIXBoxH * 4 : kIYBoxW * kIRB * 4;
                int cx, cy;
                // (measured: cp.async.bulk.prefetch.tensor boxes ahead of the loads make it slower, 0.675 -> 0.78 ms: the TMA
                // unit itself is the limiter at about one 128-byte line per 8 cycles per SM for these narrow rows)
                for (int b = 0; b < nblk; ++b, ++it) {
                    const uint32_t st = it % kIStages, ph = (it / kIStages) & 1u;
                    mbar_wait_backoff(empty0 + 8u * st, ph ^ 1u, 200);
                    origin(b, cx, cy);
                    mbar_arrive_expect_tx(full0 + 8u * st, bytes);
                    tma_load_3d(stage0 + st * (kIStageFloats * 4), map, cx, cy, d, full0 + 8u * st);
                }
            } else {
                it += nblk;
            }
            it = __shfl_sync(0xffffffffu, it, 0);
            continue;
        }

        // ---------------- consumers: thread = chain c0 + tid ----------------
        const int c = c0 + tid;
        float acc = 0.f;
        bool have = false;
        if (mode == 2) {
            // y-major: step i visits row y = i (or H-1-i), the chain is at x = c + R(i)
            const bool rev = ip.ry[d] < 0;
            for (int b = 0; b < nblk; ++b, ++it) {
                const uint32_t st = it % kIStages, ph = (it / kIStages) & 1u;
                const int i0 = b * kIRB, nk = min(dm.H - i0, kIRB);
                const int Rl = R[min(i0 + lane, dm.H - 1)];                   // lane k: shift of step i0 + k
                const int Ra = __shfl_sync(0xffffffffu, Rl, 0), Rb = __shfl_sync(0xffffffffu, Rl, nk - 1);
                const int Rmin = min(Ra, Rb);
                const float* tile = stages + (size_t)st * kIStageFloats + tid - Rmin + ((c0 + Rmin) & 3);   // + trow * kIYBoxW + R_k: this chain
                const long long ystep = rev ? -(long long)dm.pitch : (long long)dm.pitch;
                float* row = P + (long long)(rev ? dm.H - 1 - i0 : i0) * dm.pitch + c;  // + R_k: this chain's pixel of step i0
                const int tstep = rev ? -kIYBoxW : kIYBoxW;
                tile += rev ? (kIRB - 1) * kIYBoxW : 0;
                // every lane's chain inside the image for the whole tile (x = c + R is monotone along the tile)?
                const bool inside = (unsigned)(c + Ra) < (unsigned)dm.W && (unsigned)(c + Rb) < (unsigned)dm.W;
                const bool fast = __all_sync(0xffffffffu, inside && have);
                mbar_wait_backoff(full0 + 8u * st, ph, 40);
                if (fast) {
#pragma unroll 8
                    for (int k = 0; k < nk; ++k) {
                        const int Rk = __shfl_sync(0xffffffffu, Rl, k);
                        acc = tile[k * tstep + Rk] + acc;
                        row[k * ystep + Rk] = acc;
                    }
                } else {
                    for (int k = 0; k < nk; ++k) {
                        const int Rk = __shfl_sync(0xffffffffu, Rl, k);
                        if ((unsigned)(c + Rk) < (unsigned)dm.W) {
                            const float a = tile[k * tstep + Rk];
                            if (have) { acc = a + acc; row[k * ystep + Rk] = acc; }
                            else { acc = a; have = true; }
                        } else {
                            have = false;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8u * st);
            }
        } else {
            // x-major: step i visits column x = i (or W-1-i), the chain is at y = c + R(i)
            const bool fwd = !(ip.rx[d] < 0);
            const int grp = lane >> 3;                                        // this lane's lag in the column walk
            for (int b = 0; b < nblk; ++b, ++it) {
                const uint32_t st = it % kIStages, ph = (it / kIStages) & 1u;
                const int i0 = b * kIRB, ncols = min(dm.W - i0, kIRB);
                const int Rl = R[min(i0 + lane, dm.W - 1)];
                const int Ra = __shfl_sync(0xffffffffu, Rl, 0), Rb = __shfl_sync(0xffffffffu, Rl, ncols - 1);
                const int Rmin = min(Ra, Rb);
                const int offl = Rl - Rmin;                                   // tile row of chain c0 at step i0 + lane
                const int ybase = c0 + Rmin;
                float* tile = stages + (size_t)st * kIStageFloats;
                const int xb = fwd ? i0 : dm.W - kIRB - i0;                  // image column of memory column xoff
                const int xoff = xb & 3;                                      // (the box starts at xb & ~3)
                const int mb = xoff + (fwd ? 0 : kIRB - 1), ms = fwd ? 1 : -1;   // memory column of step j: mb + ms * j
                const bool interior = ncols == kIRB && ybase >= 0 && ybase + kIXBoxH <= dm.H;
                const bool fast = __all_sync(0xffffffffu, have) && interior;
                mbar_wait_backoff(full0 + 8u * st, ph, 40);
                if (fast) {
#pragma unroll 7
                    for (int s = 0; s < kIRB + 3; ++s) {
                        const int j = s - grp;
                        const int off = __shfl_sync(0xffffffffu, offl, j & 31);
                        if ((unsigned)j < (unsigned)kIRB) {
                            float* e = tile + (tid + off) * kIXBoxW + mb + ms * j;
                            acc = *e + acc;
                            *e = acc;
                        }
                    }
                } else {
                    for (int s = 0; s < kIRB + 3; ++s) {
                        const int j = s - grp;
                        const int off = __shfl_sync(0xffffffffu, offl, j & 31);
                        if ((unsigned)j < (unsigned)ncols) {
                            const int r = tid + off;
                            if ((unsigned)(ybase + r) < (unsigned)dm.H) {
                                float* e = tile + r * kIXBoxW + mb + ms * j;
                                if (have) { acc = *e + acc; *e = acc; }
                                else { acc = *e; have = true; }
                            } else {
                                have = false;
                            }
                        }
                    }
                }
                named_bar_sync(1, kICW);                                       // every chain of the tile is summed
                // ---- store: tile row r, lane = memory column; the element belongs to chain (r - off of its step) ----
                const int jl = fwd ? lane : kIRB - 1 - lane;                  // step of this lane's memory column
                const int off_m = __shfl_sync(0xffffffffu, offl, jl);
                const int x = xb + lane;
                const bool col_ok = jl < ncols;
                if (interior && xoff == 0) {   // (uniform over the CTA: the row distribution below must be the same in every warp)
                    // interior tile: rows [off_max, kICW + off_min) are owned in every column -> whole 128-byte rows as float4
                    // (8 lanes per row, 4 rows per warp instruction); the ragged rows above / below go element-wise
                    const int off_lo = Ra < Rb ? 0 : Ra - Rb, off_hi = Ra < Rb ? Rb - Ra : 0;   // off of the first / last step ...
                    const int full_lo = max(off_lo, off_hi), full_hi = kICW + min(off_lo, off_hi);   // ... off is monotone between them
                    const int sub = lane >> 3, l8 = lane & 7;
                    for (int r = full_lo + warp * 4 + sub; r < full_hi; r += kIConsumerWarps * 4) {
                        const float4 v = *reinterpret_cast<const float4*>(tile + r * kIXBoxW + 4 * l8);
                        *reinterpret_cast<float4*>(P + (long long)(ybase + r) * dm.pitch + xb + 4 * l8) = v;
                    }
                    for (int r = warp; r < kIXBoxH; r += kIConsumerWarps) {
                        if (r >= full_lo && r < full_hi) { r += ((full_hi - 1 - r) / kIConsumerWarps) * kIConsumerWarps; continue; }
                        if ((unsigned)(r - off_m) < (unsigned)kICW) P[(long long)(ybase + r) * dm.pitch + x] = tile[r * kIXBoxW + lane];
                    }
                } else {
                    float* gp = P + (long long)(ybase + warp) * dm.pitch + x;
                    const float* tp = tile + warp * kIXBoxW + lane + xoff;
#pragma unroll 4
                    for (int r = warp; r < kIXBoxH; r += kIConsumerWarps, gp += (size_t)kIConsumerWarps * dm.pitch, tp += kIConsumerWarps * kIXBoxW)
                        if (col_ok && (unsigned)(r - off_m) < (unsigned)kICW && (unsigned)(ybase + r) < (unsigned)dm.H) *gp = *tp;
                }
                fence_proxy_async();                                          // our tile writes precede the next TMA fill of this stage
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8u * st);
            }
        }
    }
}

size_t integral_tma_smem_bytes() { return (size_t)kIStages * kIStageFloats * sizeof(float); }
int integral_strip_chains() { return kICW; }

bool integral_tma_encode(const void* planes, const MapDims& dm, CUtensorMap* map_y, CUtensorMap* map_x) {
    return encode_planes_map(map_y, planes, dm.W, dm.H, dm.D, dm.pitch, kIYBoxW, kIRB, 1) &&
           encode_planes_map(map_x, planes, dm.W, dm.H, dm.D, dm.pitch, kIXBoxW, kIXBoxH, 1);
}

void launch_integral_tma(float* d_planes, const MapDims& dm, const IntegralParams& ip, const IntegralPlanDev& plan, int n_sms,
                         cudaStream_t s) {
    if (plan.n_items <= 0) return;
    const size_t smem = integral_tma_smem_bytes();
    cudaFuncSetAttribute(integral_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 2;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, integral_tma_kernel, kIThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 2;
    const int grid = plan.n_items < per_sm * n_sms ? plan.n_items : per_sm * n_sms;
    integral_tma_kernel<<<grid, kIThreads, smem, s>>>(plan.map_y, plan.map_x, d_planes, dm, ip, plan.rtab, plan.rlen, plan.items,
                                                      plan.n_items, plan.counter);
}

}   // namespace fdcm
