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
// Items carry the range of tiles in which their strip meets the image (the tiles outside are never loaded) and are handed
// out through an atomic counter in decreasing order of work (persistent CTAs, as many per SM as fit).
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

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(kIThreads, 4)
integral_tma_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x,
                    float* __restrict__ planes, MapDims dm, const __grid_constant__ IntegralParams ip,
                    const int4* __restrict__ items, int n_items,
                    int* __restrict__ counter) {
    extern __shared__ __align__(128) float stages[];          // [kIStages][kIStageFloats]
    __shared__ __align__(8) uint64_t bars[2 * kIStages];      // full[0..S), empty[S..2S)
    __shared__ int s_item;
    __shared__ int s_xtbl[kIConsumerWarps][kIRB + 8];         // x-major: per-warp table of the tile's per-step byte offsets ...
    __shared__ int s_xoff[kIConsumerWarps][kIRB + 8];         // ... and tile row offsets (entries 3 .. 3 + kIRB - 1 are the steps)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kIStages);
    const uint32_t stage0 = smem_u32(stages);
    if (tid < kIConsumerWarps * (kIRB + 8)) {                 // (the pad entries are read, never used)
        (&s_xtbl[0][0])[tid] = 0;
        (&s_xoff[0][0])[tid] = 0;
    }
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
        const int d = items[item].x, c0 = items[item].y, b_lo = items[item].z, b_hi = items[item].w;   // tiles [b_lo, b_hi) meet the image
        const int mode = ip.mode[d];
        // R(i) = (long)roundf(float(i) * r) (imgproc.h:55,72) evaluated in place: the same IEEE multiply and round-half-away
        // as the host table of the plan (which sizes the work items), and no table load on the critical path of a tile
        const float rslope = mode == 1 ? ip.ry[d] : ip.rx[d];
        auto R = [&](int i) { return (int)roundf((float)i * rslope); };
        float* P = planes + (size_t)d * dm.plane_elems;
        const int n_major = mode == 1 ? dm.W : dm.H;
        const int nblk = b_hi - b_lo;

        if (warp == kIConsumerWarps) {
            // ---------------- producer: one thread issues every box load of the item ----------------
            if (lane == 0) {
                const bool rev = (mode == 1 ? ip.rx[d] : ip.ry[d]) < 0;
                // box origin of tile b: (innermost coordinate rounded down to 16 bytes, second coordinate)
                auto origin = [&](int b, int& cx, int& cy) {
                    const int i0 = b * kIRB, i1 = min(n_major, i0 + kIRB);
                    const int minor0 = c0 + min(R(i0), R(i1 - 1));            // R is monotone in i
                    const int major0 = rev ? n_major - kIRB - i0 : i0;
                    if (mode == 1) { cx = major0 & ~3; cy = minor0; }
                    else { cx = minor0 & ~3; cy = major0; }
                };
                const CUtensorMap* map = mode == 1 ? &map_x : &map_y;
                const uint32_t bytes = mode == 1 ? kIXBoxW * kIXBoxH * 4 : kIYBoxW * kIRB * 4;
                int cx, cy;
                // (measured: cp.async.bulk.prefetch.tensor boxes ahead of the loads make it slower, 0.675 -> 0.78 ms: the TMA
                // unit itself is the limiter at about one 128-byte line per 8 cycles per SM for these narrow rows)
                for (int b = b_lo; b < b_hi; ++b, ++it) {
                    const uint32_t st = it % kIStages, ph = (it / kIStages) & 1u;
                    origin(b, cx, cy);                                        // (table loads done before the stage is free)
                    mbar_wait_backoff(empty0 + 8u * st, ph ^ 1u, 200);
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
            for (int b = b_lo; b < b_hi; ++b, ++it) {
                const uint32_t st = it % kIStages, ph = (it / kIStages) & 1u;
                const int i0 = b * kIRB, nk = min(dm.H - i0, kIRB);
                const int Rl = R(min(i0 + lane, dm.H - 1));                   // lane k: shift of step i0 + k
                const int Ra = __shfl_sync(0xffffffffu, Rl, 0), Rb = __shfl_sync(0xffffffffu, Rl, nk - 1);
                const int Rmin = min(Ra, Rb);
                const float* tile = stages + (size_t)st * kIStageFloats + tid - Rmin + ((c0 + Rmin) & 3);   // + trow * kIYBoxW + R_k: this chain
                const int ystep = rev ? -dm.pitch : dm.pitch;                 // (32-bit element offsets: one address instruction per store)
                float* row = P + (long long)(rev ? dm.H - 1 - i0 : i0) * dm.pitch + c;  // + R_k: this chain's pixel of step i0
                const int tstep = rev ? -kIYBoxW : kIYBoxW;
                tile += rev ? (kIRB - 1) * kIYBoxW : 0;
                // every lane's chain inside the image for the whole tile (x = c + R is monotone along the tile)?
                const bool inside = (unsigned)(c + Ra) < (unsigned)dm.W && (unsigned)(c + Rb) < (unsigned)dm.W;
                const bool fast = __all_sync(0xffffffffu, inside && have);
                mbar_wait_backoff(full0 + 8u * st, ph, 40);
                // (addresses as opaque bases + small offsets: one address instruction per load and two per store; the loads
                // of half a tile are issued together, then the dependent additions and the row stores)
                uint32_t ta = smem_u32(tile);
                asm volatile("" : "+r"(ta), "+l"(row));
                __builtin_assume(__isGlobal(row));
                const int tsb = rev ? -kIYBoxW * 4 : kIYBoxW * 4;
                constexpr int kHalf = kIRB / 2;
                if (fast && nk == kIRB) {
                    // whole tile, every chain inside: fully unrolled, tile offsets are immediates
#pragma unroll
                    for (int k0 = 0; k0 < kIRB; k0 += kHalf) {
                        float v[kHalf];
                        int go[kHalf];
#pragma unroll
                        for (int i = 0; i < kHalf; ++i) {
                            const int k = k0 + i;
                            const int Rk = __shfl_sync(0xffffffffu, Rl, k);
                            v[i] = lds_f32(ta + (uint32_t)(Rk << 2) + (uint32_t)(k * tsb));
                            go[i] = k * ystep + Rk;
                        }
#pragma unroll
                        for (int i = 0; i < kHalf; ++i) {
                            acc = v[i] + acc;
                            row[go[i]] = acc;
                        }
                    }
                } else {
                    // tile at the image border (chains start / end inside it) or the last, shorter tile: the same two phases
                    // with per-lane predicates instead of branches (a strip that crosses the border obliquely spends many
                    // tiles here, and its tiles are a serial chain)
#pragma unroll
                    for (int k0 = 0; k0 < kIRB; k0 += kHalf) {
                        float v[kHalf];
                        int go[kHalf];
                        unsigned in = 0u;
#pragma unroll
                        for (int i = 0; i < kHalf; ++i) {
                            const int k = k0 + i;
                            const int Rk = __shfl_sync(0xffffffffu, Rl, k);
                            const bool ok = k < nk && (unsigned)(c + Rk) < (unsigned)dm.W;
                            v[i] = 0.f;
                            if (ok) v[i] = lds_f32(ta + (uint32_t)(Rk << 2) + (uint32_t)(k * tsb));
                            go[i] = k * ystep + Rk;
                            in |= ok ? 1u << i : 0u;
                        }
#pragma unroll
                        for (int i = 0; i < kHalf; ++i) {
                            const bool ok = (in >> i) & 1u;
                            const bool st_ok = ok && have;                 // the first pixel of a chain keeps its value
                            acc = st_ok ? v[i] + acc : (ok ? v[i] : acc);
                            if (st_ok) row[go[i]] = acc;
                            if (k0 + i < nk) have = ok;
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
            for (int b = b_lo; b < b_hi; ++b, ++it) {
                const uint32_t st = it % kIStages, ph = (it / kIStages) & 1u;
                const int i0 = b * kIRB, ncols = min(dm.W - i0, kIRB);
                const int Rl = R(min(i0 + lane, dm.W - 1));
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
                    // Interior tile.  The byte offset of step j relative to the chain's own tile row is the same for every
                    // chain: lane j puts it into the warp's table.  Slot s of the walk is step s - grp (the skew), so a lane
                    // reads the table through a pointer shifted by its lag.  The walk runs in chunks: all loads of a chunk,
                    // then the dependent additions, then the stores -- a load never waits behind a store of the same chunk.
                    s_xtbl[warp][lane + 3] = (offl * kIXBoxW + mb + ms * lane) * 4;
                    __syncwarp();
                    const int* tl = &s_xtbl[warp][3 - grp];
                    char* lb = reinterpret_cast<char*>(tile) + tid * (kIXBoxW * 4);
                    constexpr int kSlots = kIRB + 3, kChunk = 12;
#pragma unroll
                    for (int s0 = 0; s0 < kSlots; s0 += kChunk) {
                        float v[kChunk];
                        float* a[kChunk];
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            const int s = s0 + i;
                            if (s < kSlots) {
                                const bool ok = (s >= 3 && s < kIRB) || (unsigned)(s - grp) < (unsigned)kIRB;
                                a[i] = reinterpret_cast<float*>(lb + tl[s]);
                                if (ok) v[i] = *a[i];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            const int s = s0 + i;
                            if (s < kSlots) {
                                const bool ok = (s >= 3 && s < kIRB) || (unsigned)(s - grp) < (unsigned)kIRB;
                                if (ok) { acc = v[i] + acc; v[i] = acc; }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            const int s = s0 + i;
                            if (s < kSlots) {
                                const bool ok = (s >= 3 && s < kIRB) || (unsigned)(s - grp) < (unsigned)kIRB;
                                if (ok) *a[i] = v[i];
                            }
                        }
                    }
                } else {
                    // Tile at the image border (chains start / end inside it) or the last, narrower tile: the same chunked
                    // walk with per-lane predicates instead of branches.  A second table holds the tile row offset of every
                    // step: chain tid is inside the image at step j iff 0 <= ybase + tid + off_j < H.
                    s_xtbl[warp][lane + 3] = (offl * kIXBoxW + mb + ms * lane) * 4;
                    s_xoff[warp][lane + 3] = offl;
                    __syncwarp();
                    const int* tl = &s_xtbl[warp][3 - grp];
                    const int* ol = &s_xoff[warp][3 - grp];
                    char* lb = reinterpret_cast<char*>(tile) + tid * (kIXBoxW * 4);
                    const int y0 = ybase + tid;
                    constexpr int kSlots = kIRB + 3, kChunk = 12;
#pragma unroll
                    for (int s0 = 0; s0 < kSlots; s0 += kChunk) {
                        float v[kChunk];
                        float* a[kChunk];
                        unsigned in = 0u, jv = 0u;
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            const int s = s0 + i;
                            if (s < kSlots) {
                                const bool jok = (unsigned)(s - grp) < (unsigned)ncols;
                                const bool ok = jok && (unsigned)(y0 + ol[s]) < (unsigned)dm.H;
                                a[i] = reinterpret_cast<float*>(lb + tl[s]);
                                v[i] = 0.f;
                                if (ok) v[i] = *a[i];
                                jv |= jok ? 1u << i : 0u;
                                in |= ok ? 1u << i : 0u;
                            }
                        }
                        unsigned stm = 0u;
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            if (s0 + i < kSlots) {
                                const bool ok = (in >> i) & 1u;
                                const bool st_ok = ok && have;             // the first pixel of a chain keeps its value
                                acc = st_ok ? v[i] + acc : (ok ? v[i] : acc);
                                v[i] = acc;
                                stm |= st_ok ? 1u << i : 0u;
                                if ((jv >> i) & 1u) have = ok;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            if (s0 + i < kSlots) {
                                if ((stm >> i) & 1u) *a[i] = v[i];
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
                    // (all shared-memory reads of a group first, then its global stores: the stores of one row do not wait
                    // behind the read of the next)
                    {
                        constexpr int kG = 4;                                    // float4 rows per lane and round (<= 8 in all)
                        const int r0 = full_lo + warp * 4 + sub;
                        float* g4 = P + (long long)(ybase + r0) * dm.pitch + xb + 4 * l8;
#pragma unroll
                        for (int h = 0; h < kICW / (kIConsumerWarps * 4 * kG); ++h) {
                            float4 v[kG];
#pragma unroll
                            for (int i = 0; i < kG; ++i) {
                                const int r = r0 + (h * kG + i) * kIConsumerWarps * 4;
                                if (r < full_hi) v[i] = *reinterpret_cast<const float4*>(tile + r * kIXBoxW + 4 * l8);
                            }
#pragma unroll
                            for (int i = 0; i < kG; ++i) {
                                const int r = r0 + (h * kG + i) * kIConsumerWarps * 4;
                                if (r < full_hi) *reinterpret_cast<float4*>(g4 + (h * kG + i) * kIConsumerWarps * 4 * dm.pitch) = v[i];
                            }
                        }
                    }
                    float* gcol = P + (long long)ybase * dm.pitch + x;
                    // ragged rows above (r < full_lo <= kIRB) and below (r >= full_hi >= kICW) the full rows: at most kIRB each
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        constexpr int kR = kIRB / kIConsumerWarps;               // 8 rows per warp and part
                        const int rb = part == 0 ? warp : full_hi + warp, re = part == 0 ? full_lo : kIXBoxH;
                        float w[kR];
#pragma unroll
                        for (int i = 0; i < kR; ++i) {
                            const int r = rb + i * kIConsumerWarps;
                            if (r < re && (unsigned)(r - off_m) < (unsigned)kICW) w[i] = tile[r * kIXBoxW + lane];
                        }
#pragma unroll
                        for (int i = 0; i < kR; ++i) {
                            const int r = rb + i * kIConsumerWarps;
                            if (r < re && (unsigned)(r - off_m) < (unsigned)kICW) gcol[r * dm.pitch] = w[i];
                        }
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
int integral_tile_steps() { return kIRB; }

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
    integral_tma_kernel<<<grid, kIThreads, smem, s>>>(plan.map_y, plan.map_x, d_planes, dm, ip, plan.items4,
                                                      plan.n_items, plan.counter);
}

}   // namespace fdcm
