// dt_band_kernels.cu — exact-regime distance transform, "band" formulation (sm_100a).
//
// Reference: core::distanceTransform (core/imgproc.h:160-195) = two calls of _distanceTransformColumnPassL2
// (imgproc.h:91-130) on a {0, FLT_MAX} image, first down the columns, then along the rows (with the in-place aliasing
// of the second loop, imgproc.h:122-128).  In the exact regime (2*(side-1)^2 < 2^24, see dt3_kernels.cu) both calls are
// integer arithmetic, so they can be restated as
//   column call: g(x, y) = distance to the nearest edge pixel of column x                       (dt_col_band_kernel)
//   row call:    owner(q) = leftmost argmin_v g(v)^2 + (q - v)^2,
//                out(q)   = (owner < q ? out(owner) : g(owner)^2) + (q - owner)^2                (dt_row_band_kernel)
//
// Mapping.  A *band* is 32 consecutive rows of one orientation plane.  dt_col_band_kernel transposes the 1-bit edge mask
// into one 8-byte record per (band, column): the 32 edge bits of the column inside the band and the distances from the
// band to the nearest edge above / below it.  From that record every row of the band derives g with two bit scans, so
// the u16 distance image of the first formulation (N/2 bytes written and read back) no longer exists.
// dt_row_band_kernel gives one warp per band, one LANE PER ROW: every lane runs the sequential lower-envelope stack
// algorithm over the columns (no idle lanes: 32 independent rows per warp, neighbouring rows take almost the same
// branches), the top of the stack in registers, the next 32 entries in a shared-memory ring, older ones spilled to a
// per-row global array.  The pixel fill walks the envelope once per row and goes through a 32x32 shared-memory
// transposition so that global stores are 128-byte row segments.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fdcm {

constexpr uint32_t kNone16 = 0xFFFFu;

// =============================================================================================
// per (plane, band, column): {edge bits of the 32 rows, (rows from the band's first row up to the last edge above) |
// (rows from the band's last row down to the first edge below) << 16}; 0xFFFF = no such edge
// =============================================================================================
// 32 x 32 bit-matrix transposition across a warp: lane r holds row word a (bit c = column c); afterwards lane c holds the
// column word (bit r = row r).  Five butterfly steps swap the off-diagonal blocks of size 16, 8, 4, 2, 1.
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t a, int lane) {
    uint32_t m = 0x0000FFFFu;                                  // bits whose column index has bit j clear
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, a, j);
        a = (lane & j) ? (a & ~m) | ((y >> j) & m) : (a & m) | ((y << j) & ~m);
        m ^= m << (j >> 1);                                    // 0x0000FFFF -> 0x00FF00FF -> 0x0F0F0F0F -> 0x33333333 -> 0x55555555
    }
    return a;
}

constexpr int kColSegs = 4;                                    // band segments walked in parallel per column

__global__ void __launch_bounds__(256) dt_col_band_kernel(const uint32_t* __restrict__ mask, MapDims dm,
                                                          uint2* __restrict__ info, int nbands) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* M = reinterpret_cast<uint32_t*>(smem_raw);                       // [nbands][64] column bit words
    uint16_t* up16 = reinterpret_cast<uint16_t*>(M + (size_t)nbands * 64);     // [nbands][64]
    __shared__ int s_last[kColSegs][64], s_first[kColSegs][64];                // per segment and column: last / first edge row (-1: none)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int d = blockIdx.y;
    const int w0 = blockIdx.x * 2;                                             // first mask word (32 columns each)
    const uint32_t* mp = mask + (size_t)d * dm.H * dm.wwords;
    // ---- transpose: 32 rows x 64 columns of mask bits -> 64 column words ----
    for (int b = warp; b < nbands; b += nwarps) {
        const int y = b * 32 + lane;
        uint32_t a0 = 0, a1 = 0;
        if (y < dm.H) {
            a0 = mp[(size_t)y * dm.wwords + w0];
            if (w0 + 1 < dm.wwords) a1 = mp[(size_t)y * dm.wwords + w0 + 1];
        }
        uint32_t c0 = 0, c1 = 0;
        if (__any_sync(0xffffffffu, (a0 | a1) != 0)) {
            c0 = warp_bit_transpose(a0, lane);
            c1 = warp_bit_transpose(a1, lane);
        }
        M[(size_t)b * 64 + lane] = c0;
        M[(size_t)b * 64 + 32 + lane] = c1;
    }
    __syncthreads();
    // ---- nearest edge above / below every band, per column: thread = (segment of bands, column) ----
    const int c = threadIdx.x & 63, seg = threadIdx.x >> 6;
    const int per = (nbands + kColSegs - 1) / kColSegs;
    const int b0 = min(nbands, seg * per), b1 = min(nbands, b0 + per);
    {
        int last = -1, first = -1;
        for (int b = b0; b < b1; ++b) {
            const uint32_t m = M[(size_t)b * 64 + c];
            if (m) {
                if (first < 0) first = b * 32 + __ffs(m) - 1;
                last = b * 32 + 31 - __clz(m);
            }
        }
        s_last[seg][c] = last;
        s_first[seg][c] = first;
    }
    __syncthreads();
    int last = -1, next = -1;                                                  // last edge row before / first edge row after the segment
    for (int sg = 0; sg < seg; ++sg) { const int v = s_last[sg][c]; if (v >= 0) last = v; }
    for (int sg = kColSegs - 1; sg > seg; --sg) { const int v = s_first[sg][c]; if (v >= 0) next = v; }
    for (int b = b0; b < b1; ++b) {
        up16[(size_t)b * 64 + c] = (uint16_t)(last < 0 ? kNone16 : (uint32_t)(b * 32 - last));
        const uint32_t m = M[(size_t)b * 64 + c];
        if (m) last = b * 32 + 31 - __clz(m);
    }
    const int x = blockIdx.x * 64 + c;
    uint2* out = info + (size_t)d * nbands * dm.pitch + x;
    for (int b = b1 - 1; b >= b0; --b) {
        const uint32_t m = M[(size_t)b * 64 + c];
        const uint32_t dd = next < 0 ? kNone16 : (uint32_t)(next - (b * 32 + 31));
        if (x < dm.pitch) out[(size_t)b * dm.pitch] = make_uint2(m, (uint32_t)up16[(size_t)b * 64 + c] | (dd << 16));
        if (m) next = b * 32 + __ffs(m) - 1;
    }
}

// =============================================================================================
// row call, one lane per row
// =============================================================================================
constexpr int kRing = 32;                       // stack entries below the top kept in shared memory, per row

// stack entry: x = f(v) + v^2, y = v | (first owned pixel) << 16
__device__ __forceinline__ void sts_u64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The envelope stack of one row: top entry in registers, the kRing entries below it in a shared-memory ring (this
// lane's column of [kRing][32] 8-byte slots), older ones in the row's global array (a sequential walk stays inside
// one 128-byte line for 16 entries, so only one access in 16 pays the L2 latency).
// Two stacks per row: the LEFT one takes the columns below the split in ascending order and stores for every vertex the
// first pixel it owns; the RIGHT one takes the columns from the split on in DESCENDING order (mirror image: stores the
// last pixel a vertex owns).  The left stack grows from the front of the row's array, the right one from its back, so
// that afterwards the array reads in ascending vertex order: [0, K_left) and [maxdepth - K_right, maxdepth).
struct RowStack {
    uint32_t ring0;         // shared address of this lane's slot 0
    uint32_t so;            // byte offset of slot (k & (kRing-1)): where entry k goes when it is displaced from the registers
    uint2* spill;           // entry k lives at spill[sdir * k]
    int sdir;
    int k, cnt;             // index of the top entry; the ring holds entries [k - cnt, k - 1]
    uint32_t topkey, topvs;
    int topv, tops;         // vertex and first (left stack) / last (right stack) owned pixel of the top

    __device__ __forceinline__ void init(uint32_t ring_addr, uint2* slot0, int dir) {
        ring0 = ring_addr;
        so = (kRing - 1) * 256u;
        spill = slot0;
        sdir = dir;
        k = -1; cnt = 0; topkey = 0; topvs = 0; topv = 0; tops = 0;
    }
    __device__ __forceinline__ void push(int v, int bound, uint32_t key) {
        if (k >= 0) {
            if (cnt == kRing) spill[sdir * (k - kRing)] = lds_u64(ring0 + so);   // ring full: its oldest entry (same slot) leaves
            else ++cnt;
            sts_u64(ring0 + so, make_uint2(topkey, topvs));
        }
        ++k;
        so = (so + 256u) & (kRing * 256u - 1u);
        topv = v; tops = bound; topkey = key;
        topvs = (uint32_t)v | ((uint32_t)bound << 16);
    }
    __device__ __forceinline__ void pop() {
        --k;
        so = (so - 256u) & (kRing * 256u - 1u);
        if (k < 0) return;
        uint2 e;
        if (cnt > 0) {
            --cnt;
            e = lds_u64(ring0 + so);
        } else {
            e = spill[sdir * k];
        }
        topkey = e.x;
        topvs = e.y;
        topv = (int)(e.y & 0xFFFFu);
        tops = (int)(e.y >> 16);
    }
    // floor(N / Dn) for 0 <= N / Dn < 2897: the biased approximate quotient is below the true one by less than 0.003
    static __device__ __forceinline__ int floor_div(int N, int Dn) {
        const float qf = __fmaf_rn((float)N, rcp_approx((float)Dn), -0.002f);
        const int t = __float2int_rd(qf);        // floor(N / Dn) or one less
        return t + ((N - t * Dn) >= Dn ? 1 : 0);
    }
    // first pixel from which parabola (v, key), v right of every vertex of the stack, beats the envelope (imgproc.h:
    // 104-120 on integers; the left parabola keeps ties); pops the vertices that end up owning nothing; >= W: never
    // kLoose (candidate pass, see dt_row_band_kernel): the top is only popped when the new parabola takes over at least one
    // pixel BEFORE the top's first one; a vertex popped that way is nowhere strictly minimal, even between pixels
    template <bool kLoose = false>
    __device__ __forceinline__ int take_over(int v, uint32_t key, int Wm1) {
        while (k >= 0) {
            const int N = (int)key - (int)topkey;
            const int Dn = 2 * (v - topv);
            if (N < (tops - (kLoose ? 1 : 0)) * Dn) {   // from a pixel not after the top's first one: the top owns nothing
                pop();
                continue;
            }
            if (N >= Wm1 * Dn) return Wm1 + 1;   // beyond the last pixel of the row
            return floor_div(N, Dn) + 1;
        }
        return 0;
    }
    // left stack, one column: ascending v
    template <bool kLoose = false>
    __device__ __forceinline__ void column(int v, int gv, int Wm1) {
        const uint32_t key = (uint32_t)(gv * gv) + (uint32_t)(v * v);
        const int start = take_over<kLoose>(v, key, Wm1);
        if (start <= Wm1) push(v, start, key);
    }
    // right stack, one column: descending v.  Parabola v (left of every vertex of the stack) is not worse than the top up
    // to pixel floor(N / Dn) (it keeps ties); the top is popped when that reaches the top's last pixel.
    template <bool kLoose = false>
    __device__ __forceinline__ void column_rev(int v, int gv, int Wm1) {
        const uint32_t key = (uint32_t)(gv * gv) + (uint32_t)(v * v);
        int end = Wm1;
        while (k >= 0) {
            const int N = (int)topkey - (int)key;
            const int Dn = 2 * (topv - v);
            if (N >= (tops + (kLoose ? 1 : 0)) * Dn) {
                pop();
                continue;
            }
            if (N < 0) return;                   // not even pixel 0: v never owns anything
            end = floor_div(N, Dn);
            break;
        }
        push(v, end, key);
    }
    // write the ring and the top to the global array; returns the number of entries
    __device__ __forceinline__ int park() {
        if (k >= 0) {
            for (int i = k - cnt; i < k; ++i) spill[sdir * i] = lds_u64(ring0 + (uint32_t)(i & (kRing - 1)) * 256u);
            spill[sdir * k] = make_uint2(topkey, topvs);
        }
        return k + 1;
    }
};

// per-row result of the envelope kernel
struct RowMeta {
    int32_t k_left;         // entries [0, k_left) of the row array, {f(v) + v^2, v | first owned pixel << 16}
    int32_t k_right;        // entries [maxdepth - k_right, maxdepth), {f(v) + v^2, v | LAST owned pixel << 16}
    int32_t right_start;    // first pixel of the first right entry (the following ones start after their predecessor's last)
    int32_t pad;
};

// kFromG = false: g is derived from the band records of dt_col_band_kernel (product path).
// kFromG = true : g rows are given explicitly (u16, 0xFFFF = FLT_MAX), any content (row-pass parity tests).
// One CTA of two warps per band, lane = row: warp 0 builds the left stack over the columns [0, xsplit), warp 1 the right
// stack over [xsplit, W) from the right; warp 0 then joins them: the right entries are offered to the left stack in
// ascending order until one keeps a pixel (the lower envelopes of two column ranges cross exactly once).  The serial
// chain per row is half as long as with one stack and twice as many warps are in flight.
// Workspace rows are padded to 32 per band (row id = (d * nbands + b) * 32 + lane) so that the rows past H of the last
// band need no special case: they build an envelope nobody reads.
// cand (shared memory, may be null): 1 bit per column, columns whose bit is clear are skipped.
template <bool kFromG, bool kRev>
__device__ __forceinline__ void envelope_half(RowStack& st, const uint2* __restrict__ info_row, const uint16_t* __restrict__ g,
                                              const uint16_t* __restrict__ g_row, const MapDims& dm, int d, int row0, int x_begin,
                                              int x_end, int lane, int rsel, const uint32_t* __restrict__ cand, int* claim = nullptr) {
    const int W = dm.W, Wm1 = dm.W - 1;
    const uint32_t mle = 0xFFFFFFFFu >> (31 - rsel), mge = 0xFFFFFFFFu << rsel;
    const int l31 = 31 - rsel;
    const uint2 kNone = make_uint2(0u, 0xFFFFFFFFu);
    const int nchunks = (x_end - x_begin) >> 5;
    auto chunk_x0 = [&](int c) { return kRev ? x_end - 32 * (c + 1) : x_begin + 32 * c; };
    uint2 e_next = kNone;
    if (!kFromG && nchunks > 0 && chunk_x0(0) + lane < W) e_next = info_row[chunk_x0(0)];
    uint32_t cw_next = 0xFFFFFFFFu;
    if (cand && nchunks > 0) cw_next = cand[chunk_x0(0) >> 5];
    for (int c = 0; c < nchunks; ++c) {
        if (claim) {
            // the two warps of a band walk towards each other over the same range and claim their chunks one by one: they meet
            // where their work (not their column count) is equal, and neither waits for the other at the join
            int ok = 0;
            if (lane == 0) ok = atomicAdd(claim, 1) < nchunks;
            if (!__shfl_sync(0xffffffffu, ok, 0)) break;
        }
        const int x0 = chunk_x0(c);
        const uint2 e = e_next;
        const uint32_t cw = cw_next;
        if (cand && c + 1 < nchunks) cw_next = cand[chunk_x0(c + 1) >> 5];
        unsigned todo;
        if (kFromG) {
            // lane = column here: which of the 32 columns hold a finite value in ANY row of the band
            unsigned any = 0;
            for (int r = 0; r < 32 && row0 + r < dm.H; ++r)
                if (g[((size_t)d * dm.H + row0 + r) * dm.pitch + x0 + lane] != kNone16) any = 1;
            todo = __ballot_sync(0xffffffffu, any && x0 + lane < W);
        } else {
            e_next = kNone;
            if (c + 1 < nchunks && chunk_x0(c + 1) + lane < W) e_next = info_row[chunk_x0(c + 1)];   // one iteration ahead
            todo = __ballot_sync(0xffffffffu, e.x != 0u || e.y != 0xFFFFFFFFu) & cw;
        }
        const bool no_bits = kFromG || __ballot_sync(0xffffffffu, e.x != 0u) == 0u;   // no edge pixel inside the band here
        while (todo) {
            const int j = kRev ? 31 - __clz(todo) : __ffs(todo) - 1;
            todo &= ~(1u << j);
            int gv;
            bool fin = true;
            if (kFromG) {
                gv = g_row[x0 + j];
                fin = row0 + lane < dm.H && gv != (int)kNone16;
            } else if (no_bits) {
                // g = distance to the nearest edge above / below the band
                const uint32_t ud = __shfl_sync(0xffffffffu, e.y, j);
                gv = min(rsel + (int)(ud & 0xFFFFu), l31 + (int)(ud >> 16));
            } else {
                const uint32_t M = __shfl_sync(0xffffffffu, e.x, j), ud = __shfl_sync(0xffffffffu, e.y, j);
                const uint32_t above = M & mle, below = M & mge;
                const int ua = __clz(above) - 31, ub = (int)(ud & 0xFFFFu);       // lane - (row of the last edge at or above)
                const int da = __ffs(below) - 1 - 31, db = (int)(ud >> 16);       // (row of the first edge at or below) - 31
                gv = min(rsel + (above ? ua : ub), l31 + (below ? da : db));
            }
            if (fin) {
                if (kRev) st.column_rev(x0 + j, gv, Wm1);
                else st.column(x0 + j, gv, Wm1);
            }
        }
    }
}

// Candidate pruning, per band.  An edge pixel P above the band that owns a pixel A of one of the band's rows is the
// STRICTLY nearest edge pixel on the whole open segment from A to P (Voronoi cells are star-shaped around their site),
// and that segment crosses the band's first row: on that row P's column has an interval of positive length on the
// continuous lower envelope (P is also that column's vertical site there, or a nearer one in the column would beat it at
// the crossing point).  The same holds for edge pixels below the band and the band's last row.  Hence
//     columns with an edge pixel inside the band
//   + survivors of a stack pass over the band's first row that only pops vertices which are nowhere strictly minimal
//   + survivors of the same pass over its last row
// is a superset of the owners of every row of the band, and all other columns can be skipped by the 32-row pass.
// The two single-row passes are themselves parallel: lane = column segment, no join (a vertex of the row's envelope is
// also a vertex of the envelope of its own segment); each lane remembers only its last kCandRing entries, older ones
// stay candidates (tests/test_envelope_model.py::test_per_band_candidates_cover_every_row_of_the_band is the CPU model).
constexpr int kCandRing = 8;

struct LooseRing {
    uint32_t ring0;         // shared address of this lane's slot 0 ([kCandRing][32] 8-byte slots, lane-interleaved)
    int head, cnt;          // the ring holds cnt entries below the top, oldest in slot head
    bool has;
    uint32_t topkey;
    int topv, tops;

    __device__ __forceinline__ void init(uint32_t addr) { ring0 = addr; head = 0; cnt = 0; has = false; topkey = 0; topv = 0; tops = 0; }
    static __device__ __forceinline__ void mark(uint32_t* cand, int v) { atomicOr(cand + (v >> 5), 1u << (v & 31)); }
    __device__ __forceinline__ void column(int v, int gv, int Wm1, uint32_t* cand) {
        const uint32_t key = (uint32_t)(gv * gv) + (uint32_t)(v * v);
        int start = 0;
        while (has) {
            const int N = (int)key - (int)topkey;
            const int Dn = 2 * (v - topv);
            if (N < (tops - 1) * Dn) {            // the top is nowhere strictly minimal (see RowStack::take_over<kLoose>)
                if (cnt > 0) {
                    --cnt;
                    const uint2 e = lds_u64(ring0 + (uint32_t)((head + cnt) & (kCandRing - 1)) * 256u);
                    topkey = e.x;
                    topv = (int)(e.y & 0xFFFFu);
                    tops = (int)(e.y >> 16);
                } else {
                    has = false;                  // nothing remembered below: the next vertex starts at pixel 0
                }
                continue;
            }
            if (N >= Wm1 * Dn) return;            // takes over beyond the last pixel: owns nothing
            start = RowStack::floor_div(N, Dn) + 1;
            break;
        }
        if (has) {
            if (cnt == kCandRing) {               // the oldest remembered entry leaves: it stays a candidate
                mark(cand, (int)(lds_u64(ring0 + (uint32_t)head * 256u).y & 0xFFFFu));
                head = (head + 1) & (kCandRing - 1);
                --cnt;
            }
            sts_u64(ring0 + (uint32_t)((head + cnt) & (kCandRing - 1)) * 256u, make_uint2(topkey, (uint32_t)topv | ((uint32_t)tops << 16)));
            ++cnt;
        }
        has = true;
        topkey = key; topv = v; tops = start;
    }
    __device__ __forceinline__ void flush(uint32_t* cand) {
        if (has) mark(cand, topv);
        for (int i = 0; i < cnt; ++i) mark(cand, (int)(lds_u64(ring0 + (uint32_t)((head + i) & (kCandRing - 1)) * 256u).y & 0xFFFFu));
    }
};

// shared memory of dt_row_band_kernel: [ring of the 32-row pass | aliased by the candidate phase: g of the band's first and
// last row (u16 per column) + the loose rings] followed by the candidate column mask
__host__ __device__ inline size_t band_alias_bytes(int pitch) {
    const size_t ring = (size_t)2 * kRing * 32 * sizeof(uint2);
    const size_t cand = (size_t)2 * pitch * sizeof(uint16_t) + (size_t)2 * kCandRing * 32 * sizeof(uint2);
    return (ring > cand ? ring : cand + 15) / 16 * 16;
}

template <bool kFromG>
__global__ void __launch_bounds__(64) dt_row_band_kernel(const uint2* __restrict__ info, const uint16_t* __restrict__ g, MapDims dm,
                                                         int nbands, uint2* __restrict__ spill_all, int maxdepth,
                                                         RowMeta* __restrict__ row_meta, int xsplit, int band_lo, int band_hi, int band_off) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    __shared__ int s_kright[32];
    __shared__ int s_claim;                                 // chunks of the 32-row pass handed out so far (both warps)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint2* ring_all = reinterpret_cast<uint2*>(band_smem);                                   // [2][kRing * 32]
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(band_smem + band_alias_bytes(dm.pitch)); // [wwords]
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring_all + (size_t)warp * kRing * 32) + (uint32_t)lane * 8u;
    const int Wm1 = dm.W - 1;

    // plane-fastest (neighbouring CTAs: different planes); the bands that overlap the scene's rows come first in the grid:
    // they carry the long serial chains
    // band_off: first band (in that order) of this launch -- the scene bands and the far bands can be launched separately
    const int d = blockIdx.x % dm.D;
    int b = band_off + blockIdx.x / dm.D;
    {
        const int n_in = band_hi - band_lo + 1;
        b = b < n_in ? band_lo + b : (b - n_in < band_lo ? b - n_in : b);
    }
    const int row0 = b * 32;
    const uint2* info_row = info + ((size_t)d * nbands + b) * dm.pitch + lane;
    const uint16_t* g_row = kFromG ? g + ((size_t)d * dm.H + min(row0 + lane, dm.H - 1)) * dm.pitch : nullptr;
    const size_t prow = ((size_t)d * nbands + b) * 32 + lane;
    uint2* row_entries = spill_all + prow * maxdepth;

    const uint32_t* cand = nullptr;
    if (threadIdx.x == 0) s_claim = 0;                      // (the candidate phase below has two CTA barriers before its first use)
    if (!kFromG) {
        // ---- candidate phase ----
        // A band inside the rows that can hold edge pixels needs both passes: warp 0 takes the band's first row, warp 1 its
        // last (existing) row.  A band above all of them (b < band_lo) only needs its last row, a band below all of them its
        // first row (every owner of its rows is an edge pixel on that one side): there the two warps share that single row,
        // half of the columns each -- half the latency and half the instructions of the pass, and no survivors of a row
        // that proves nothing.
        const bool one_sided = b < band_lo || b > band_hi;     // (uniform over the CTA; the launcher widens [band_lo, band_hi] when unsure)
        uint16_t* s_g = reinterpret_cast<uint16_t*>(band_smem) + (one_sided ? (size_t)0 : (size_t)warp * dm.pitch);
        const int rsel = one_sided ? (b < band_lo ? min(31, dm.H - 1 - row0) : 0) : (warp == 0 ? 0 : min(31, dm.H - 1 - row0));
        const uint32_t mle = 0xFFFFFFFFu >> (31 - rsel), mge = 0xFFFFFFFFu << rsel;
        bool any = false;
        // (four record loads in flight per lane: the loop runs at the latency of its loads otherwise -- 12 % of the kernel's
        // stall samples sat on the first use of `e`)
        constexpr int kStageUnroll = 4;
        const int xstep = one_sided ? 64 : 32;
        for (int xb = one_sided ? warp * 32 : 0; xb < dm.pitch; xb += kStageUnroll * xstep) {
            uint2 ev[kStageUnroll];
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int x0 = xb + u * xstep;
                ev[u] = (x0 < dm.pitch && x0 + lane < dm.W) ? info_row[x0] : make_uint2(0u, 0xFFFFFFFFu);
            }
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int x0 = xb + u * xstep;
                if (x0 >= dm.pitch) break;                 // (uniform)
                const uint2 e = ev[u];
                uint32_t gv = 0xFFFFu;
                if (e.x != 0u || e.y != 0xFFFFFFFFu) {
                    const uint32_t above = e.x & mle, below = e.x & mge;
                    const int up = rsel + (above ? __clz(above) - 31 : (int)(e.y & 0xFFFFu));
                    const int dn = (31 - rsel) + (below ? __ffs(below) - 1 - 31 : (int)(e.y >> 16));
                    gv = (uint32_t)min(up, dn);
                    any = true;
                }
                s_g[x0 + lane] = (uint16_t)gv;
                const uint32_t inside = __ballot_sync(0xffffffffu, e.x != 0u);
                if ((one_sided || warp == 0) && lane == 0) s_cand[x0 >> 5] = inside;   // columns with an edge pixel inside the band
            }
        }
        __syncthreads();
        // lane = column segment; the segment length in 16-bit words is 2 (mod 4): consecutive lanes hit different banks
        const int nseg = one_sided ? 64 : 32;
        int seg = (dm.W + nseg - 1) / nseg;
        seg += (2 - (seg & 3)) & 3;
        LooseRing lr;
        lr.init((uint32_t)__cvta_generic_to_shared(band_smem + (size_t)2 * dm.pitch * sizeof(uint16_t) +
                                                   (size_t)warp * kCandRing * 32 * sizeof(uint2)) + (uint32_t)lane * 8u);
        if (one_sided || __any_sync(0xffffffffu, any)) {
            const int v0 = (one_sided ? warp * 32 + lane : lane) * seg, v1 = min(dm.W, v0 + seg);
            for (int v = v0; v < v1; ++v) {
                const int gv = s_g[v];
                if (gv != 0xFFFF) lr.column(v, gv, Wm1, s_cand);
            }
            lr.flush(s_cand);
        }
        __syncthreads();                                    // the mask is complete; the staging area becomes the ring
        cand = s_cand;
    }

    // product path: no fixed split column, the warps claim 32-column chunks from their side until they meet (s_claim);
    // explicit-g path (tests): the launcher's split column
    int* claim = kFromG ? nullptr : &s_claim;
    RowStack st;
    if (warp == 1) {
        st.init(ring_addr, row_entries + (maxdepth - 1), -1);
        envelope_half<kFromG, true>(st, info_row, g, g_row, dm, d, row0, kFromG ? xsplit : 0, dm.pitch, lane, lane, cand, claim);
        s_kright[lane] = st.park();
    } else {
        st.init(ring_addr, row_entries, 1);
        envelope_half<kFromG, false>(st, info_row, g, g_row, dm, d, row0, 0, kFromG ? xsplit : dm.pitch, lane, lane, cand, claim);
    }
    __syncthreads();
    if (warp != 0) return;
    // ---- join: offer the right entries (ascending vertex) to the left stack until one keeps a pixel ----
    const int k_right = s_kright[lane];
    const uint2* right = row_entries + (maxdepth - k_right);
    int j = 0, right_start = 0;
    while (j < k_right) {
        const uint2 r = right[j];
        ++j;
        const int v = (int)(r.y & 0xFFFFu), end = (int)(r.y >> 16);
        const int start = st.take_over(v, r.x, Wm1);
        if (start <= end) {                      // it owns [start, end]: it becomes the top of the left stack
            st.push(v, start, r.x);
            right_start = end + 1;
            break;
        }
    }
    RowMeta meta;
    meta.k_left = st.park();
    meta.k_right = k_right - j;
    meta.right_start = right_start;
    meta.pad = 0;
    row_meta[prow] = meta;
}

// =============================================================================================
// fill (imgproc.h:122-128 incl. its in-place aliasing): out(q) = base(owner(q)) + (q - v_owner)^2, lane = pixel of a
// 32-pixel chunk.  The row's envelope (vertex v, first owned pixel s, f(v) + v^2; s strictly increasing) is read 32
// entries at a time, one entry per lane.
//   base: the reference's second loop writes out(q) = img(v) + (q-v)^2 while img(v) may already hold out(v).  A vertex
//   inside its own interval (s <= v) contributes f(v); a vertex left of it (s > v) contributes out(v) = base(j) +
//   (v - v_j)^2 with j the entry that owns pixel v (the last one with s_j <= v, normally one of the nearest few below).
//   Bases are resolved once per batch, in registers: backward search by shuffles over the current and the previous
//   batch, then propagation along the (downward pointing) dependencies; a global-memory walk covers the rest.
//   fill: the entries whose interval starts inside the chunk set one bit each in `marks`; the owner of pixel p is the
//   popc(marks & bits <= p)-th of them (or the entry carried over from the left).
// =============================================================================================
struct RowFill {
    const uint2* row;       // the row's array: left entries at the front, right entries at the back (see RowMeta)
    int KL, roff, rstart;   // left entries; array slot of right entry i is i + roff; first pixel of the first right entry
    int K, e0, ci, nxt;     // entries; first entry of the batch; entries consumed; first pixel of entry ci (uniform)
    uint2 be, nbe, pbe;     // batch entry of this lane (x = resolved base), same lane of the next (raw) / previous (resolved) batch
    uint32_t raw_last;      // stored second field (first / last pixel) of the last lane of the current batch (uniform)
    int carry_v;            // the entry that owns the pixel left of the current chunk (uniform)
    uint32_t carry_b;

    // entry i of the joined envelope as {f(v) + v^2, v | first owned pixel << 16}
    __device__ __forceinline__ uint2 entry(int i) const {
        if (i < KL) return row[i];
        uint2 e = row[i + roff];
        const uint32_t start = i == KL ? (uint32_t)rstart : (row[i + roff - 1].y >> 16) + 1u;   // predecessor's last pixel + 1
        e.y = (e.y & 0xFFFFu) | (start << 16);
        return e;
    }
    // batch loading: one raw load per lane; the right entries get their first pixel from the lane below (the last pixel
    // of the predecessor, + 1), lane 0 from the last lane of the previous batch
    __device__ __forceinline__ uint2 load_raw(int i) const { return i < K ? row[i < KL ? i : i + roff] : make_uint2(0u, 0xFFFFFFFFu); }
    __device__ __forceinline__ uint2 convert(uint2 raw, int i, int lane, uint32_t prev_last) const {
        const uint32_t below = __shfl_up_sync(0xffffffffu, raw.y >> 16, 1);
        if (i >= KL && i < K) {
            const uint32_t start = i == KL ? (uint32_t)rstart : (lane ? below : prev_last) + 1u;
            raw.y = (raw.y & 0xFFFFu) | (start << 16);
        }
        return raw;
    }
    __device__ uint32_t slow_base(int idx) const {
        uint32_t acc = 0;
        int cur = idx;
        while (true) {
            const uint2 e = entry(cur);
            const int v = (int)(e.y & 0xFFFFu), s = (int)(e.y >> 16);
            if (s <= v) return e.x - (uint32_t)(v * v) + acc;
            int lo = 0, hi = cur - 1;            // largest o with s_o <= v (s_0 = 0)
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((int)(entry(mid).y >> 16) <= v) lo = mid; else hi = mid - 1;
            }
            const int dd = v - (int)(entry(lo).y & 0xFFFFu);
            acc += (uint32_t)(dd * dd);
            cur = lo;
        }
    }
    // kResolvedBelow: the entries before the previous batch have already been replaced by their resolved form {base,
    // v | first pixel << 16} (the fused kernel's resolve pass writes every batch back before it moves on)
    __device__ uint32_t slow_base_resolved(int v) const {
        int lo = 0, hi = e0 - 33;                // largest o <= e0 - 33 with first pixel <= v (entry 0 starts at pixel 0)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((int)(row[mid < KL ? mid : mid + roff].y >> 16) <= v) lo = mid; else hi = mid - 1;
        }
        const uint2 e = row[lo < KL ? lo : lo + roff];
        const int dd = v - (int)(e.y & 0xFFFFu);
        return e.x + (uint32_t)(dd * dd);
    }
    template <bool kResolvedBelow = false>
    __device__ __forceinline__ void resolve(int lane) {
        const int bv = (int)(be.y & 0xFFFFu), bs = (int)(be.y >> 16);
        uint32_t base = be.x - (uint32_t)(bv * bv);
        const bool chain = e0 + lane < K && bs > bv;
        if (__any_sync(0xffffffffu, chain)) {
            // owner = the last entry with s <= v: s increases along the lanes, so five doubling steps find the last lane
            // below this one that qualifies; the previous batch is searched the same way when there is none
            int owner = -1;                      // lane of the owner entry; once found, negative = lane owner + 32 of the previous batch
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int t = owner + step;
                const uint32_t y = __shfl_sync(0xffffffffu, be.y, t & 31);
                if (t < lane && (int)(y >> 16) <= bv) owner = t;
            }
            bool found = !chain || owner >= 0;
            if (__any_sync(0xffffffffu, !found)) {
                int po = -1;
#pragma unroll
                for (int step = 32; step >= 1; step >>= 1) {            // lanes 0..31 (po + 32 = 31 first)
                    const int t = po + step;
                    const uint32_t y = __shfl_sync(0xffffffffu, pbe.y, t & 31);
                    if (t <= 31 && (int)(y >> 16) <= bv) po = t;
                }
                if (!found && po >= 0) {
                    owner = po - 32;
                    found = true;
                }
            }
            bool ready = !chain;
            if (chain && !found) {               // owner further than 32 entries below: walk the global array
                base = kResolvedBelow ? slow_base_resolved(bv) : slow_base(e0 + lane);
                ready = true;
            }
            const int src = owner & 31;
            {
                const uint32_t pb = __shfl_sync(0xffffffffu, pbe.x, src);
                const int pv = (int)(__shfl_sync(0xffffffffu, pbe.y, src) & 0xFFFFu);
                if (!ready && chain && owner < 0) {
                    const int dd = bv - pv;
                    base = pb + (uint32_t)(dd * dd);
                    ready = true;
                }
            }
            unsigned rdy = __ballot_sync(0xffffffffu, ready);
            while (rdy != 0xFFFFFFFFu) {         // owners inside the batch: each round resolves at least the lowest pending lane
                const uint32_t ob = __shfl_sync(0xffffffffu, base, src);
                const int ov = __shfl_sync(0xffffffffu, bv, src);
                if (!ready && ((rdy >> src) & 1u)) {
                    const int dd = bv - ov;
                    base = ob + (uint32_t)(dd * dd);
                    ready = true;
                }
                rdy = __ballot_sync(0xffffffffu, ready);
            }
        }
        be.x = base;
    }
    __device__ __forceinline__ void init(const uint2* row_entries, const RowMeta meta, int maxdepth, int lane) {
        row = row_entries;
        KL = meta.k_left;
        roff = maxdepth - meta.k_right - meta.k_left;
        rstart = meta.right_start;
        K = meta.k_left + meta.k_right; e0 = 0; ci = 0; nxt = K > 0 ? 0 : 0x7FFFFFFF;
        const uint2 raw = load_raw(lane);    // (sentinel s = 0xFFFF: never starts inside a chunk)
        raw_last = __shfl_sync(0xffffffffu, raw.y >> 16, 31);
        be = convert(raw, lane, lane, 0u);
        nbe = load_raw(32 + lane);           // raw: converted when it becomes the current batch
        pbe = make_uint2(0u, 0xFFFFFFFFu);
        carry_v = 0; carry_b = 0;
        resolve(lane);
    }
    // value of pixel q0 + lane; chunks must be visited left to right
    __device__ __forceinline__ uint32_t chunk(int q0, int lane, uint32_t le_mask) {
        int ov = carry_v;
        uint32_t ob = carry_b;
        while (nxt < q0 + 32) {                               // some interval starts inside this chunk
            if (ci == e0 + 32) {                              // ... in the next 32 entries
                e0 += 32;
                pbe = be;
                const uint32_t prev_last = raw_last;
                raw_last = __shfl_sync(0xffffffffu, nbe.y >> 16, 31);
                be = convert(nbe, e0 + lane, lane, prev_last);
                nbe = load_raw(e0 + 32 + lane);
                resolve(lane);
            }
            const int bv = (int)(be.y & 0xFFFFu);
            const int rel = (int)(be.y >> 16) - q0;
            const bool inb = (unsigned)rel < 32u;             // consecutive lanes starting at ci - e0 (s is increasing)
            const unsigned bal = __ballot_sync(0xffffffffu, inb);
            const unsigned marks = __reduce_or_sync(0xffffffffu, inb ? (1u << rel) : 0u);
            const int c = __popc(marks & le_mask);
            const int src = (ci - e0 + c - 1) & 31;
            const uint32_t sb = __shfl_sync(0xffffffffu, be.x, src);
            const int sv = __shfl_sync(0xffffffffu, bv, src);
            if (c) { ov = sv; ob = sb; }
            const int lastl = 31 - __clz(bal);
            carry_v = __shfl_sync(0xffffffffu, bv, lastl);
            carry_b = __shfl_sync(0xffffffffu, be.x, lastl);
            ci += __popc(bal);
            if (ci >= K) nxt = 0x7FFFFFFF;
            else if (ci < e0 + 32) nxt = (int)(__shfl_sync(0xffffffffu, be.y, ci - e0) >> 16);
            else if (ci < KL) nxt = (int)(__shfl_sync(0xffffffffu, nbe.y, 0) >> 16);      // first entry of the next (raw) batch
            else nxt = ci == KL ? rstart : (int)raw_last + 1;
        }
        const int dq = q0 + lane - ov;
        return ob + (uint32_t)(dq * dq);
    }

    // ---- batch-driven fill (fused kernel) ----
    // Inside the scene's rows most envelope entries own one to four pixels (every edge pixel of a line that runs along the
    // row is a site of its own), so walking the pixels and looking the owners up pays the bookkeeping per 32 pixels many
    // times per batch of entries, and one plane with a thousand entries holds its warp (and the CTA's barrier) back.
    // The fused kernel first runs a resolve pass (warp = plane, batches in order, previous batch in registers): it turns
    // every entry into {base, v | first pixel << 16} and writes it back in place.  After that a batch is a self-contained
    // unit of work, and the batches of all planes are dealt to the warps: a plane with a thousand entries in this row no
    // longer holds one warp back while the others wait at the barrier.
    int bend;               // first pixel of the entry that follows the batch (uniform); 0xFFFF: none
    // resolve pass (sequential over the batches of one plane's row): write the current batch back in resolved form
    __device__ __forceinline__ void write_back(int lane) const {
        const int i = e0 + lane;
        if (i < K) const_cast<uint2*>(row)[i < KL ? i : i + roff] = be;
    }
    // (the pass does little per batch, so it runs at the latency of its loads: three batches are kept in flight; a batch
    // is written back only after its successors were loaded, so everything this function converts is still raw)
    __device__ __forceinline__ void advance_resolved(int lane, uint2& nbe2, uint2& nbe3) {
        e0 += 32;
        pbe = be;
        const uint32_t prev_last = raw_last;
        raw_last = __shfl_sync(0xffffffffu, nbe.y >> 16, 31);
        be = convert(nbe, e0 + lane, lane, prev_last);
        nbe = nbe2;
        nbe2 = nbe3;
        nbe3 = load_raw(e0 + 96 + lane);
        resolve<true>(lane);
    }
    // a batch of the resolved array as a unit of the fill
    __device__ __forceinline__ void load_unit(const uint2* row_entries, int kl, int k, int roff_, int j, int bend_, int lane) {
        row = row_entries;
        KL = kl; K = k; roff = roff_;
        e0 = 32 * j;
        bend = bend_;
        be = load_raw(e0 + lane);
    }
    // writes the pixels of the batch inside [q0, q1) into trow (the tile row of this plane, pixel q0 at trow[0]) as float
    // bits of the (exact) squared distances.  The batch owns the pixels [first pixel of its entry 0, bend); they are
    // taken 32 at a time, lane = pixel: the owner of pixel p is the last entry that starts at or before p = (entries
    // starting at or before the piece's first pixel) + (entries starting inside the piece up to p) - 1.  Pieces are
    // independent of each other (no carry, no look-ahead), four are in flight per round.
    __device__ __forceinline__ void write_unit(int q0, int q1, uint32_t* __restrict__ trow, int lane) const {
        const int s = (int)(be.y >> 16);                          // (sentinel lanes past the last entry: 0xFFFF, never counted)
        const int v = (int)(be.y & 0xFFFFu);
        const uint32_t base = be.x;
        const int p_end = min(bend, q1);
        const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);
        auto piece = [&](int p0) {
            const int rel = s - p0;
            const int nb = __popc(__ballot_sync(0xffffffffu, rel <= 0));
            const unsigned marks = __reduce_or_sync(0xffffffffu, (unsigned)(rel - 1) < 31u ? 1u << rel : 0u);
            const int idx = nb - 1 + __popc(marks & le_mask);
            const int ov = __shfl_sync(0xffffffffu, v, idx);
            const uint32_t ob = __shfl_sync(0xffffffffu, base, idx);
            const int p = p0 + lane;
            const int dq = p - ov;
            if (p < p_end) trow[p - q0] = __float_as_uint((float)(ob + (uint32_t)(dq * dq)));
        };
        int p0 = max(__shfl_sync(0xffffffffu, s, 0), q0);
        for (; p0 + 96 < p_end; p0 += 128) {
            piece(p0);
            piece(p0 + 32);
            piece(p0 + 64);
            piece(p0 + 96);
        }
        for (; p0 < p_end; p0 += 32) piece(p0);
    }
};

constexpr int kFillWarps = 4;

__device__ __forceinline__ size_t padded_row(int row, int H) {   // [D*H] row id -> workspace row (32 rows per band)
    return (size_t)(row / H) * (size_t)(((H + 31) >> 5) << 5) + (size_t)(row % H);
}

// stand-alone fill (stage-wise builds, depths without a fused kernel): one warp per row
__global__ void __launch_bounds__(kFillWarps * 32) dt_row_fill_kernel(const uint2* __restrict__ spill_all,
                                                                      const RowMeta* __restrict__ row_meta,
                                                                      float* __restrict__ planes, MapDims dm, int n_rows_total,
                                                                      int maxdepth) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kFillWarps + warp;           // row of the [D*H][pitch] stack of planes
    if (row >= n_rows_total) return;
    const int W = dm.W;
    float* op = planes + (size_t)row * dm.pitch + lane;       // this lane's pixel of the current chunk
    const size_t prow = padded_row(row, dm.H);
    const RowMeta meta = row_meta[prow];
    if (meta.k_left + meta.k_right <= 0) {                    // no edge pixel in this plane: FLT_MAX stays (imgproc.h:174)
        for (int q = lane; q < W; q += 32, op += 32) *op = FLT_MAX;
        return;
    }
    RowFill rf;
    rf.init(spill_all + prow * maxdepth, meta, maxdepth, lane);
    const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);
    for (int q0 = 0; q0 < dm.pitch; q0 += 32, op += 32) {
        const uint32_t val = rf.chunk(q0, lane, le_mask);
        if (q0 + lane < W) *op = (float)val;                  // < 2^24: exact
    }
}

// =============================================================================================
// fused fill + propagateOrientation (matching/src/featuremaps/dt3cpu.cpp:77-107): one CTA per image row y.
// Warp w fills the pixels [q0, q0 + kFPChunk) of its planes' row y into a shared-memory tile [D][kFPChunk]; then every
// thread takes one pixel, runs the circular min-plus sweeps over its D values in registers (with the sqrt of the L2
// transform, core/imgproc.h:191-192, applied on the way in) and writes the D results as coalesced row segments.  The
// distance-transform planes are never written to or read back from HBM (2N bytes less than fill -> propagate).
// =============================================================================================
template <int D>
struct FPConfig {
    static constexpr int kPlanesPerWarp = 2;
    static constexpr int kWarps = (D + kPlanesPerWarp - 1) / kPlanesPerWarp;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kChunk = kThreads;             // pixels per iteration = one per thread in the propagate phase
};

// sqrtf for the values the fill phase produces: integers 0 .. 2^24 (exactly representable squared distances) and FLT_MAX
// (no edge in the plane).  rsqrt.approx + one Newton step carried out with two fused multiply-adds is correctly rounded
// for every one of them (fdcm_debug_sqrt_check compares all 2^24 + 2 inputs with the IEEE sqrtf on the device;
// tests/test_gpu_parity.py::test_fast_sqrt_is_exact); 6 instructions instead of the ~10 of the general routine, which
// also has to handle subnormals and non-integers.  0 * inf = NaN for the input 0 is mapped back by the final max.
__device__ __forceinline__ float sqrt_exact_int(float n) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n));
    const float s = __fmul_rn(n, r);
    const float h = __fmul_rn(0.5f, r);
    const float e = __fmaf_rn(-s, s, n);
    return fmaxf(__fmaf_rn(e, h, s), 0.f);
}

__global__ void sqrt_check_kernel(unsigned long long* __restrict__ n_bad, uint32_t* __restrict__ first_bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. 2^24 + 1
    float x = (float)i;
    if (i == (1u << 24) + 1u) x = FLT_MAX;
    if (i > (1u << 24) + 1u) return;
    if (__float_as_uint(sqrt_exact_int(x)) != __float_as_uint(sqrtf(x))) {
        if (atomicAdd(n_bad, 1ull) == 0ull) *first_bad = i;
    }
}

// number of inputs (integers 0 .. 2^24 and FLT_MAX) on which sqrt_exact_int differs from the IEEE sqrtf
cudaError_t run_sqrt_check(unsigned long long* h_bad, uint32_t* h_first, cudaStream_t s) {
    unsigned long long* d_bad = nullptr;
    cudaError_t e = cudaMalloc(&d_bad, 16);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(d_bad, 0, 16, s);
    sqrt_check_kernel<<<((1u << 24) + 2u + 255u) / 256u, 256, 0, s>>>(d_bad, reinterpret_cast<uint32_t*>(d_bad + 1));
    unsigned long long h[2] = {0, 0};
    e = cudaMemcpyAsync(h, d_bad, 16, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_bad);
    *h_bad = h[0];
    *h_first = (uint32_t)h[1];
    return e;
}

// propagate phase shared by the fused kernels: thread t takes pixel q0 + t of image row y, reads its D values from the
// tile (float bits written by the fill phase), runs the forward (ceil(1.5 D)) and backward (D + floor(1.5 D)) circular sweeps
// P[c2] = min(P[c2], P[c1] + w_step) in registers and writes D coalesced row segments.
template <int D>
__device__ __forceinline__ void propagate_from_tile(const uint32_t* tile, int chunk, int q0, int y, float* __restrict__ planes,
                                                    const MapDims& dm, const PropParams& pp, int sqrt_first) {
    const int x = q0 + (int)threadIdx.x;
    if (x >= dm.W) return;
    float v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        v[d] = __uint_as_float(tile[(size_t)d * chunk + threadIdx.x]);   // the fill phase stored float bits (FLT_MAX: no edge)
    }
    if (sqrt_first) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = sqrt_exact_int(v[d]);
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
    float* op = planes + (size_t)y * dm.pitch + x;
#pragma unroll
    for (int d = 0; d < D; ++d) op[(size_t)d * dm.plane_elems] = v[d];
}

// Resolve pass of the fused fill: one warp per (plane, row) walks the row's envelope batch by batch (previous batch in
// registers, three batches of loads in flight), turns every entry into {base, v | first pixel << 16} -- right-stack entries
// get their first pixel, chained bases are resolved -- and writes it back in place.  After this kernel every batch of 32
// entries is a self-contained unit of work for dt_fill_propagate_kernel.  Rows [ya0, ya0 + na) then [yb0, ...).
constexpr int kResolveWarps = 8;
__global__ void __launch_bounds__(kResolveWarps * 32)
dt_resolve_kernel(uint2* spill_all /* rewritten in place: no __restrict__, coherent loads */, const RowMeta* __restrict__ row_meta,
                  MapDims dm, int maxdepth, int ya0, int na, int yb0, int n_rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * kResolveWarps + warp;             // plane-fastest: neighbouring warps, different planes of one row
    if (w >= n_rows * dm.D) return;
    const int ri = w / dm.D, d = w - ri * dm.D;
    const int y = ri < na ? ya0 + ri : yb0 + (ri - na);
    const int Hp = ((dm.H + 31) >> 5) << 5;
    const size_t prow = (size_t)d * Hp + y;
    const RowMeta meta = row_meta[prow];
    const int K = meta.k_left + meta.k_right;
    if (K <= 0) return;
    RowFill rf;
    rf.init(spill_all + prow * maxdepth, meta, maxdepth, lane);
    uint2 nbe2 = rf.load_raw(64 + lane), nbe3 = rf.load_raw(96 + lane);
    for (;;) {
        rf.write_back(lane);
        if (rf.e0 + 32 >= K) break;
        __syncwarp();                                            // (later batches may look entries of this one up: slow_base_resolved)
        rf.advance_resolved(lane, nbe2, nbe3);
    }
}

constexpr int kFPMaxBatches = 96;        // envelope entries per row <= window width <= 2897 -> at most 91 batches of 32
constexpr int kFPMaxChunks = 8;

template <int D>
__global__ void __launch_bounds__(FPConfig<D>::kThreads, 3)
dt_fill_propagate_kernel(const uint2* __restrict__ spill_all /* resolved entries (dt_resolve_kernel) */,
                         const RowMeta* __restrict__ row_meta, float* __restrict__ planes,
                         MapDims dm, int maxdepth, const __grid_constant__ PropParams pp, int sqrt_first, int ya0, int na, int yb0) {
    using C = FPConfig<D>;
    extern __shared__ __align__(16) uint32_t fp_tile[];      // [D][kChunk] squared distances as float bits (exact: < 2^24; FLT_MAX: no edge)
    __shared__ uint16_t s_first[D][kFPMaxBatches + 1];       // first pixel of every batch of every plane's envelope (then 0xFFFF)
    __shared__ int4 s_meta[D];                               // {k_left, entries, array offset of the right entries, first right pixel}
    __shared__ uint16_t s_pref[kFPMaxChunks][32];            // per chunk: units of the planes before plane d ([D]: all units)
    __shared__ uint8_t s_jl[kFPMaxChunks][32];               // per chunk: first batch of plane d that reaches into the chunk
    __shared__ int s_next_unit;                              // units of the current chunk handed out so far
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = (int)blockIdx.x < na ? ya0 + (int)blockIdx.x : yb0 + ((int)blockIdx.x - na);   // rows [ya0, ya0 + na) then [yb0, ...)
    const int Hp = ((dm.H + 31) >> 5) << 5;                  // workspace rows per plane
    const int nchunks = (dm.W + C::kChunk - 1) / C::kChunk;

    if (threadIdx.x == 0) s_next_unit = 0;
    // ---- first pixel of every batch (warp w: planes w and w + kWarps; the arrays hold resolved entries: dt_resolve_kernel) ----
#pragma unroll
    for (int p = 0; p < C::kPlanesPerWarp; ++p) {
        const int d = warp + p * C::kWarps;
        if (d < D) {
            const size_t prow = (size_t)d * Hp + y;
            const RowMeta meta = row_meta[prow];
            const int K = meta.k_left + meta.k_right, roff = maxdepth - K;
            if (lane == 0) s_meta[d] = make_int4(meta.k_left, K, roff, meta.right_start);
            const uint2* row = spill_all + prow * maxdepth;
            const int nb = (K + 31) >> 5;
            for (int j = lane; j <= kFPMaxBatches; j += 32) {
                const int i = 32 * j;
                s_first[d][j] = (uint16_t)(j < nb ? row[i < meta.k_left ? i : i + roff].y >> 16 : 0xFFFFu);
            }
        }
    }
    __syncthreads();
    // ---- units per chunk (warp c: chunk c, lane = plane): batches jl .. jh reach into [q0, q1) ----
    if (warp < nchunks && warp < kFPMaxChunks) {
        const int q0 = warp * C::kChunk, q1 = q0 + C::kChunk;
        int n = 0, jl = 0;
        if (lane < D) {
            const int nb = (s_meta[lane].y + 31) >> 5;
            if (nb > 0) {
                int lo = 0, hi = nb - 1;                     // smallest j with first[j + 1] > q0 (first[nb] = 0xFFFF)
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)s_first[lane][mid + 1] > q0) hi = mid; else lo = mid + 1; }
                jl = lo;
                lo = jl; hi = nb - 1;                        // largest j with first[j] < q1 (first[jl] < q1: rows start at pixel 0)
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)s_first[lane][mid] < q1) lo = mid; else hi = mid - 1; }
                n = (int)s_first[lane][jl] < q1 ? lo - jl + 1 : 0;
            }
        }
        int pre = n;                                         // inclusive prefix sum over the planes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
        s_pref[warp][lane] = (uint16_t)(pre - n);            // lanes >= D carry the total
        s_jl[warp][lane] = (uint8_t)jl;
    }
    __syncthreads();

    for (int ch = 0; ch < nchunks; ++ch) {
        const int q0 = ch * C::kChunk;
        // ---- fill: the batches that reach into the chunk, claimed one by one by the warps ----
        {
            const int pref = s_pref[ch][lane], jl = s_jl[ch][lane];     // lane = plane
            const int n_units = __shfl_sync(0xffffffffu, pref, D);
            {
                // (the entries of the next unit are loaded before the current one is written: the L2 round trip of a unit
                // hides behind the pixel stores of its predecessor)
                auto load = [&](int u, RowFill& rf) -> int {
                    const unsigned le = __ballot_sync(0xffffffffu, lane < D && pref <= u);
                    const int d = 31 - __clz(le);            // the last plane whose units start at or before u
                    const int j = __shfl_sync(0xffffffffu, jl, d) + u - __shfl_sync(0xffffffffu, pref, d);
                    const int4 m = s_meta[d];
                    rf.load_unit(spill_all + ((size_t)d * Hp + y) * maxdepth, m.x, m.y, m.z, j, (int)s_first[d][j + 1], lane);
                    return d;
                };
                // (units are claimed, not dealt: a warp that drew batches of one-pixel entries takes fewer of them, and the
                // warps reach the barrier behind the fill together)
                auto claim = [&]() {
                    int u = 0;
                    if (lane == 0) u = atomicAdd(&s_next_unit, 1);
                    return __shfl_sync(0xffffffffu, u, 0);
                };
                RowFill cur, nxt;
                int u = claim(), d_cur = 0, d_nxt = 0;
                if (u < n_units) d_cur = load(u, cur);
                while (u < n_units) {
                    const int u2 = claim();
                    if (u2 < n_units) d_nxt = load(u2, nxt);
                    cur.write_unit(q0, q0 + C::kChunk, fp_tile + (size_t)d_cur * C::kChunk, lane);
                    cur.be = nxt.be;
                    cur.bend = nxt.bend;
                    d_cur = d_nxt;
                    u = u2;
                }
            }
            // planes without an edge pixel: FLT_MAX stays (imgproc.h:174)
#pragma unroll
            for (int p = 0; p < C::kPlanesPerWarp; ++p) {
                const int d = warp + p * C::kWarps;
                if (d < D && s_meta[d].y == 0) {
                    uint32_t* trow = fp_tile + (size_t)d * C::kChunk;
                    for (int c = lane; c < C::kChunk; c += 32) trow[c] = __float_as_uint(FLT_MAX);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_next_unit = 0;               // (nobody claims between the two barriers)
        // ---- propagate: one pixel per thread ----
        propagate_from_tile<D>(fp_tile, C::kChunk, q0, y, planes, dm, pp, sqrt_first);
        __syncthreads();
    }
}

// =============================================================================================
// L1 transform (core/imgproc.h:137-146,178-184): the second (row) call is two min-plus sweeps, on integers
// out(x) = min(x + min_{v<=x}(g(v) - v), -x + min_{v>=x}(g(v) + v)), g = vertical distance from the band records.
// A lane owns FOUR consecutive columns (two 16-byte loads of band records), a warp step covers 128 columns: the
// prefix / suffix minima run serially over the lane's four values and only the lane totals go through the warp scan
// (10 shuffles per 128 pixels instead of per 32).  A first right-to-left pass leaves the suffix minimum of g(v) + v per
// 128-column chunk in shared memory; the fill then walks left to right with the prefix carried across chunks.
// Exact at any map size; no envelope, no workspace.  (Columns in [W, pitch) hold "no edge" records: dt_col_band_kernel
// writes every column of the pitch and the mask has no bit there.)
// =============================================================================================
constexpr int kBigL1 = 1 << 28;
constexpr int kL1Cols = 128;                     // columns per warp step

struct L1Fill {
    const uint2* info;      // band records of (plane, band of the row), this lane's first column of chunk 0: + 4 * lane
    int* S;                 // shared: S[c] = min over columns >= 128 c of g(v) + v, S[nchunks] = big
    int r, pitch, carry;
    uint32_t mle, mge;

    // records of the four columns x0 .. x0 + 3, x0 = q0 + 4 * lane (info already points at this lane's column of chunk 0)
    __device__ __forceinline__ void load(int q0, int x0, uint4& a, uint4& b) const {
        a = b = make_uint4(0u, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
        if (x0 < pitch) {                                      // (pitch is a multiple of 32: all four columns or none)
            const uint4* p = reinterpret_cast<const uint4*>(info + q0);
            a = p[0];
            b = p[1];
        }
    }
    __device__ __forceinline__ int g_of(uint32_t ex, uint32_t ey) const {   // vertical distance of this row in the record's column, or big
        if (ex == 0u && ey == 0xFFFFFFFFu) return kBigL1;
        const uint32_t above = ex & mle, below = ex & mge;
        const int upd = r + (above ? __clz(above) - 31 : (int)(ey & 0xFFFFu));
        const int dnd = (31 - r) + (below ? __ffs(below) - 1 - 31 : (int)(ey >> 16));
        return min(upd, dnd);
    }
    __device__ __forceinline__ void bind(const uint2* info_row, int row_in_band, int pitch_, int* s_suffix) {
        info = info_row;
        S = s_suffix;
        r = row_in_band;
        pitch = pitch_;
        carry = kBigL1;
        mle = 0xFFFFFFFFu >> (31 - r);
        mge = 0xFFFFFFFFu << r;
    }
    // right-to-left pass: suffix minima of g(v) + v per chunk
    __device__ __forceinline__ void init(int lane) {
        const int nch = (pitch + kL1Cols - 1) / kL1Cols;
        int run = kBigL1;
        if (lane == 0) S[nch] = kBigL1;
        uint4 a, b;
        load((nch - 1) * kL1Cols, (nch - 1) * kL1Cols + 4 * lane, a, b);
        for (int c = nch - 1; c >= 0; --c) {
            const uint4 ca = a, cb = b;
            if (c > 0) load((c - 1) * kL1Cols, (c - 1) * kL1Cols + 4 * lane, a, b);      // one chunk ahead
            const int x0 = c * kL1Cols + 4 * lane;
            const int g0 = g_of(ca.x, ca.y), g1 = g_of(ca.z, ca.w), g2 = g_of(cb.x, cb.y), g3 = g_of(cb.z, cb.w);
            int m = g0 >= kBigL1 ? kBigL1 : g0 + x0;
            m = min(m, g1 >= kBigL1 ? kBigL1 : g1 + x0 + 1);
            m = min(m, g2 >= kBigL1 ? kBigL1 : g2 + x0 + 2);
            m = min(m, g3 >= kBigL1 ? kBigL1 : g3 + x0 + 3);
            run = min(run, __reduce_min_sync(0xffffffffu, m));
            if (lane == 0) S[c] = run;
        }
        __syncwarp();
    }
    // values of the pixels q0 + 4 * lane .. + 3 (0xFFFFFFFF: FLT_MAX); chunks must be visited left to right
    __device__ __forceinline__ void chunk(int q0, int lane, uint32_t out[4]) {
        uint4 ea, eb;
        load(q0, q0 + 4 * lane, ea, eb);
        chunk(q0, lane, ea, eb, out);
    }
    // same with the records of the four columns already loaded
    __device__ __forceinline__ void chunk(int q0, int lane, const uint4 ea, const uint4 eb, uint32_t out[4]) {
        const int x0 = q0 + 4 * lane;
        int g[4] = {g_of(ea.x, ea.y), g_of(ea.z, ea.w), g_of(eb.x, eb.y), g_of(eb.z, eb.w)};
        int a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = g[i] >= kBigL1 ? kBigL1 : g[i] - (x0 + i);
            b[i] = g[i] >= kBigL1 ? kBigL1 : g[i] + (x0 + i);
        }
        a[1] = min(a[1], a[0]); a[2] = min(a[2], a[1]); a[3] = min(a[3], a[2]);     // prefix minima inside the lane
        b[2] = min(b[2], b[3]); b[1] = min(b[1], b[2]); b[0] = min(b[0], b[1]);     // suffix minima inside the lane
        int pa = a[3], pb = b[0];                                                    // lane totals -> inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, pa, o), tb = __shfl_down_sync(0xffffffffu, pb, o);
            if (lane >= o) pa = min(pa, ta);
            if (lane + o < 32) pb = min(pb, tb);
        }
        int xa = __shfl_up_sync(0xffffffffu, pa, 1), xb = __shfl_down_sync(0xffffffffu, pb, 1);   // exclusive: the lanes before / after
        xa = min(lane == 0 ? kBigL1 : xa, carry);
        xb = min(lane == 31 ? kBigL1 : xb, S[min(q0 / kL1Cols + 1, (pitch + kL1Cols - 1) / kL1Cols)]);   // (steps beyond the pitch: padding)
        carry = min(carry, __shfl_sync(0xffffffffu, pa, 31));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int v = min(min(a[i], xa) + (x0 + i), min(b[i], xb) - (x0 + i));
            out[i] = v >= (kBigL1 >> 1) ? 0xFFFFFFFFu : (uint32_t)v;
        }
    }
};

__device__ __forceinline__ float4 l1_as_floats(const uint32_t u[4]) {
    return make_float4(u[0] == 0xFFFFFFFFu ? FLT_MAX : (float)u[0], u[1] == 0xFFFFFFFFu ? FLT_MAX : (float)u[1],
                       u[2] == 0xFFFFFFFFu ? FLT_MAX : (float)u[2], u[3] == 0xFFFFFFFFu ? FLT_MAX : (float)u[3]);
}

// stand-alone L1 row call: one warp per row
__global__ void __launch_bounds__(kFillWarps * 32) dt_row_l1_band_kernel(const uint2* __restrict__ info, float* __restrict__ planes,
                                                                         MapDims dm, int nbands, int n_rows_total) {
    extern __shared__ __align__(16) int l1_suffix[];          // [kFillWarps][nchunks + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kFillWarps + warp;
    if (row >= n_rows_total) return;
    const int d = row / dm.H, y = row % dm.H;
    const int nch = (dm.pitch + kL1Cols - 1) / kL1Cols;
    L1Fill lf;
    lf.bind(info + ((size_t)d * nbands + (y >> 5)) * dm.pitch + 4 * lane, y & 31, dm.pitch, l1_suffix + warp * (nch + 1));
    lf.init(lane);
    float* op = planes + (size_t)row * dm.pitch + 4 * lane;
    for (int q0 = 0; q0 < dm.pitch; q0 += kL1Cols, op += kL1Cols) {
        uint32_t u[4];
        lf.chunk(q0, lane, u);
        const int x0 = q0 + 4 * lane;
        const float4 f = l1_as_floats(u);
        if (x0 + 3 < dm.W) {
            *reinterpret_cast<float4*>(op) = f;
        } else {                                             // the last columns of the row: nothing is written beyond W
            if (x0 < dm.W) op[0] = f.x;
            if (x0 + 1 < dm.W) op[1] = f.y;
            if (x0 + 2 < dm.W) op[2] = f.z;
        }
    }
}

// fused L1 row call + propagateOrientation: CTA per image row, 512-pixel chunks.  Fill: warp w < 15 takes planes w and
// w + 15, four 128-column steps per plane and chunk, one 16-byte tile store per lane and step.  Propagate: one pixel per
// thread (propagate_from_tile).
template <int D>
struct L1Config {
    static constexpr int kThreads = 512;
    static constexpr int kChunk = kThreads;
    static constexpr int kFillWarps = (D + 1) / 2;           // two planes per filling warp
    static_assert(kFillWarps * 32 <= kThreads, "");
};

template <int D>
__global__ void __launch_bounds__(L1Config<D>::kThreads, 3)
dt_l1_propagate_kernel(const uint2* __restrict__ info, float* __restrict__ planes, MapDims dm, int nbands,
                       const __grid_constant__ PropParams pp) {
    using C = L1Config<D>;
    extern __shared__ __align__(16) uint32_t fp_tile[];      // [D][kChunk] distances, then [D][nchunks + 1] suffix minima
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = blockIdx.x;
    const int nch = (dm.pitch + kL1Cols - 1) / kL1Cols;
    int* suffix = reinterpret_cast<int*>(fp_tile + (size_t)D * C::kChunk);
    // (per-plane state is the carried prefix minimum only; everything else is recomputed from d)
    int carry[2] = {kBigL1, kBigL1};
    const bool filler = warp < C::kFillWarps;
    auto plane_fill = [&](int d) {
        L1Fill lf;
        lf.bind(info + ((size_t)d * nbands + (y >> 5)) * dm.pitch + 4 * lane, y & 31, dm.pitch, suffix + d * (nch + 1));
        return lf;
    };
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int d = warp + p * C::kFillWarps;
        if (filler && d < D) plane_fill(d).init(lane);
    }
    for (int q0 = 0; q0 < dm.W; q0 += C::kChunk) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int d = warp + p * C::kFillWarps;
            if (filler && d < D) {
                L1Fill lf = plane_fill(d);
                lf.carry = carry[p];
                float4* trow = reinterpret_cast<float4*>(fp_tile + (size_t)d * C::kChunk) + lane;
                // (the records of the next step are loaded before the current one is computed; beyond the pitch load() yields
                // "no edge" records and the step writes FLT_MAX into tile columns nobody reads)
                uint4 ra, rb;
                lf.load(q0, q0 + 4 * lane, ra, rb);
#pragma unroll
                for (int c = 0; c < C::kChunk; c += kL1Cols) {
                    const uint4 ca = ra, cb = rb;
                    if (c + kL1Cols < C::kChunk) lf.load(q0 + c + kL1Cols, q0 + c + kL1Cols + 4 * lane, ra, rb);
                    uint32_t u[4];
                    lf.chunk(q0 + c, lane, ca, cb, u);
                    trow[c / 4] = l1_as_floats(u);
                }
                carry[p] = lf.carry;
            }
        }
        __syncthreads();
        propagate_from_tile<D>(fp_tile, C::kChunk, q0, y, planes, dm, pp, 0);
        __syncthreads();
    }
}

static inline unsigned cdiv_u(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

int dt_band_count(const MapDims& dm) { return (dm.H + 31) / 32; }
size_t dt_band_info_bytes(const MapDims& dm) { return (size_t)dm.D * dt_band_count(dm) * dm.pitch * sizeof(uint2); }
static size_t padded_rows(const MapDims& dm) { return (size_t)dm.D * dt_band_count(dm) * 32; }
static size_t row_k_bytes(const MapDims& dm) { return (padded_rows(dm) * sizeof(RowMeta) + 255) / 256 * 256; }
// workspace of the row call: per-row RowMeta + per-row envelope array of maxdepth entries
size_t dt_band_spill_bytes(const MapDims& dm, int maxdepth) {
    return row_k_bytes(dm) + padded_rows(dm) * (size_t)maxdepth * sizeof(uint2);
}

void launch_dt_col_band(const uint32_t* d_mask, const MapDims& dm, void* d_info, cudaStream_t s) {
    const int nbands = dt_band_count(dm);
    const size_t smem = (size_t)nbands * 64 * 6;
    dim3 grid((dm.wwords + 1) / 2, dm.D);
    // (set per call: the attribute is per device and a process may drive several devices)
    if (smem > 48 * 1024) cudaFuncSetAttribute(dt_col_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dt_col_band_kernel<<<grid, 256, smem, s>>>(d_mask, dm, reinterpret_cast<uint2*>(d_info), nbands);
}

// [win_lo, win_hi]: columns that can hold an edge pixel (envelope vertices only exist there)
struct RowWs { RowMeta* row_k; uint2* spill; int win_lo, maxdepth; };
static RowWs row_ws(const MapDims& dm, void* d_ws, int win_lo, int win_hi) {
    win_lo = win_lo < 0 ? 0 : win_lo;
    win_hi = win_hi >= dm.W ? dm.W - 1 : win_hi;
    if (win_hi < win_lo) { win_lo = 0; win_hi = dm.W - 1; }
    return RowWs{reinterpret_cast<RowMeta*>(d_ws), reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(d_ws) + row_k_bytes(dm)),
                 win_lo, win_hi - win_lo + 1};
}

// which: 0 = every band, 1 = only the bands that overlap the scene rows [row_lo, row_hi], 2 = only the others
void launch_dt_row_envelope(const void* d_info, const uint16_t* d_g, const MapDims& dm, void* d_ws, int win_lo, int win_hi,
                            int row_lo, int row_hi, int which, cudaStream_t s) {
    const RowWs ws = row_ws(dm, d_ws, win_lo, win_hi);
    const int nbands = dt_band_count(dm);
    // split column: middle of the window that can hold edge pixels, on a 32-column boundary
    const int xsplit = min(dm.pitch, max(0, ((ws.win_lo + ws.win_lo + ws.maxdepth) / 2) & ~31));
    int band_lo = (row_lo < 0 ? 0 : row_lo) >> 5, band_hi = (row_hi >= dm.H ? dm.H - 1 : row_hi) >> 5;
    if (band_hi < band_lo || band_hi >= nbands) { band_lo = 0; band_hi = nbands - 1; }
    const size_t smem = band_alias_bytes(dm.pitch) + (size_t)dm.wwords * sizeof(uint32_t);
    const int n_in = band_hi - band_lo + 1;
    const int band_off = which == 2 ? n_in : 0, n_sel = which == 0 ? nbands : (which == 1 ? n_in : nbands - n_in);
    if (n_sel <= 0) return;
    if (d_g) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(dt_row_band_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dt_row_band_kernel<true><<<(unsigned)(dm.D * n_sel), 64, smem, s>>>(nullptr, d_g, dm, nbands, ws.spill, ws.maxdepth, ws.row_k, xsplit,
                                                                            band_lo, band_hi, band_off);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(dt_row_band_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dt_row_band_kernel<false><<<(unsigned)(dm.D * n_sel), 64, smem, s>>>(reinterpret_cast<const uint2*>(d_info), nullptr, dm, nbands,
                                                                             ws.spill, ws.maxdepth, ws.row_k, xsplit, band_lo, band_hi, band_off);
    }
}

// rows of the bands that overlap [row_lo, row_hi] (same clamping as launch_dt_row_envelope): [*y0, *y1)
void dt_band_scene_rows(const MapDims& dm, int row_lo, int row_hi, int* y0, int* y1) {
    const int nbands = dt_band_count(dm);
    int band_lo = (row_lo < 0 ? 0 : row_lo) >> 5, band_hi = (row_hi >= dm.H ? dm.H - 1 : row_hi) >> 5;
    if (band_hi < band_lo || band_hi >= nbands) { band_lo = 0; band_hi = nbands - 1; }
    *y0 = band_lo * 32;
    *y1 = min(dm.H, (band_hi + 1) * 32);
}

void launch_dt_row_fill(float* d_planes, const MapDims& dm, void* d_ws, int win_lo, int win_hi, cudaStream_t s) {
    const RowWs ws = row_ws(dm, d_ws, win_lo, win_hi);
    const int rows = dm.D * dm.H;
    dt_row_fill_kernel<<<cdiv_u(rows, kFillWarps), kFillWarps * 32, 0, s>>>(ws.spill, ws.row_k, d_planes, dm, rows, ws.maxdepth);
}

bool dt_fill_propagate_supported(const MapDims& dm) { return dm.D == 30; }

// rows [ya0, ya1) and [yb0, yb1) of the image (pass 0, H, 0, 0 for every row); must precede launch_dt_fill_propagate
void launch_dt_resolve(const MapDims& dm, void* d_ws, int win_lo, int win_hi, int ya0, int ya1, int yb0, int yb1, cudaStream_t s) {
    const RowWs ws = row_ws(dm, d_ws, win_lo, win_hi);
    const int na = max(0, ya1 - ya0), nb = max(0, yb1 - yb0);
    if (na + nb <= 0) return;
    const long long warps = (long long)(na + nb) * dm.D;
    dt_resolve_kernel<<<(unsigned)((warps + kResolveWarps - 1) / kResolveWarps), kResolveWarps * 32, 0, s>>>(ws.spill, ws.row_k, dm, ws.maxdepth,
                                                                                                        ya0, na, yb0, na + nb);
}

// rows [ya0, ya1) and [yb0, yb1) of the image (pass 0, H, 0, 0 for every row)
void launch_dt_fill_propagate(float* d_planes, const MapDims& dm, void* d_ws, int win_lo, int win_hi, const PropParams& pp,
                              bool sqrt_first, int ya0, int ya1, int yb0, int yb1, cudaStream_t s) {
    const RowWs ws = row_ws(dm, d_ws, win_lo, win_hi);
    using C = FPConfig<30>;
    const size_t smem = (size_t)30 * C::kChunk * sizeof(uint32_t);
    const int na = max(0, ya1 - ya0), nb = max(0, yb1 - yb0);
    if (na + nb <= 0) return;
    cudaFuncSetAttribute(dt_fill_propagate_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dt_fill_propagate_kernel<30><<<na + nb, C::kThreads, smem, s>>>(ws.spill, ws.row_k, d_planes, dm, ws.maxdepth, pp, sqrt_first ? 1 : 0, ya0, na,
                                                                    yb0);
}

void launch_dt_row_l1_band(const void* d_info, float* d_planes, const MapDims& dm, cudaStream_t s) {
    const int nbands = dt_band_count(dm), rows = dm.D * dm.H;
    const size_t smem = (size_t)kFillWarps * ((dm.pitch + kL1Cols - 1) / kL1Cols + 1) * sizeof(int);
    dt_row_l1_band_kernel<<<cdiv_u(rows, kFillWarps), kFillWarps * 32, smem, s>>>(reinterpret_cast<const uint2*>(d_info), d_planes, dm,
                                                                                 nbands, rows);
}

void launch_dt_l1_propagate(const void* d_info, float* d_planes, const MapDims& dm, const PropParams& pp, cudaStream_t s) {
    using C = L1Config<30>;
    const size_t smem = (size_t)30 * C::kChunk * sizeof(uint32_t) + (size_t)30 * ((dm.pitch + kL1Cols - 1) / kL1Cols + 1) * sizeof(int);
    cudaFuncSetAttribute(dt_l1_propagate_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dt_l1_propagate_kernel<30><<<dm.H, C::kThreads, smem, s>>>(reinterpret_cast<const uint2*>(d_info), d_planes, dm, dt_band_count(dm), pp);
}

// shared memory the band kernels need for this map (column transposition, fused L1 tile); callers fall back to the
// first-generation kernels when it exceeds what one CTA can have
size_t dt_band_smem_bytes(const MapDims& dm) {
    const size_t col = (size_t)dt_band_count(dm) * 64 * 6;
    const size_t l1 = (size_t)30 * L1Config<30>::kChunk * sizeof(uint32_t) + (size_t)30 * ((dm.pitch + kL1Cols - 1) / kL1Cols + 1) * sizeof(int);
    const size_t env = band_alias_bytes(dm.pitch) + (size_t)dm.wwords * sizeof(uint32_t);
    return std::max(std::max(col, l1), env);
}

}   // namespace fdcm
