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
#include "common.cuh"
#include "kernels.h"

namespace fdcm {

constexpr uint32_t kNone16 = 0xFFFFu;

// =============================================================================================
// per (plane, band, column): {edge bits of the 32 rows, (rows from the band's first row up to the last edge above) |
// (rows from the band's last row down to the first edge below) << 16}; 0xFFFF = no such edge
// =============================================================================================
__global__ void __launch_bounds__(256) dt_col_band_kernel(const uint32_t* __restrict__ mask, MapDims dm,
                                                          uint2* __restrict__ info, int nbands) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* M = reinterpret_cast<uint32_t*>(smem_raw);                       // [nbands][64] column bit words
    uint16_t* up16 = reinterpret_cast<uint16_t*>(M + (size_t)nbands * 64);     // [nbands][64]
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
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const uint32_t m0 = __ballot_sync(0xffffffffu, (a0 >> c) & 1u);
                const uint32_t m1 = __ballot_sync(0xffffffffu, (a1 >> c) & 1u);
                if (lane == c) { c0 = m0; c1 = m1; }
            }
        }
        M[(size_t)b * 64 + lane] = c0;
        M[(size_t)b * 64 + 32 + lane] = c1;
    }
    __syncthreads();
    // ---- nearest edge above / below every band, per column ----
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        const int x = blockIdx.x * 64 + c;
        int last = -1;
        for (int b = 0; b < nbands; ++b) {
            up16[(size_t)b * 64 + c] = (uint16_t)(last < 0 ? kNone16 : (uint32_t)(b * 32 - last));
            const uint32_t m = M[(size_t)b * 64 + c];
            if (m) last = b * 32 + 31 - __clz(m);
        }
        int next = -1;
        uint2* out = info + (size_t)d * nbands * dm.pitch + x;
        for (int b = nbands - 1; b >= 0; --b) {
            const uint32_t m = M[(size_t)b * 64 + c];
            const uint32_t dd = next < 0 ? kNone16 : (uint32_t)(next - (b * 32 + 31));
            if (x < dm.pitch) out[(size_t)b * dm.pitch] = make_uint2(m, (uint32_t)up16[(size_t)b * 64 + c] | (dd << 16));
            if (m) next = b * 32 + __ffs(m) - 1;
        }
    }
}

// =============================================================================================
// row call, one lane per row
// =============================================================================================
constexpr int kRing = 32;                       // stack entries below the top kept in shared memory, per row
constexpr int kTileP = 33;                      // pitch of the 32x32 output transposition tile
constexpr int kBandWarps = 2;                   // warps (= bands) per CTA

// stack entry: x = f(v) + v^2 during the envelope build, then the chained base value of the vertex;
//              y = v | (first owned pixel) << 16
struct RowStack {
    uint2* ring;            // shared: [kRing][32], this lane's column at + lane
    uint2* spill;           // global: this row's own [maxdepth] array (sequential walks stay inside a 128-byte line for
                            // 16 entries, so only one access in 16 pays the L2 latency)
    int k, lo;              // index of the top entry (registers); the ring holds entries [lo, k-1]
    uint32_t topkey;
    int topv, tops;
    __device__ __forceinline__ void push(int v, int start, uint32_t key) {
        if (k >= 0) {
            const int slot = (k & (kRing - 1)) * 32;
            if (k - lo == kRing) {               // ring full: its oldest entry (same slot) moves to global memory
                spill[lo] = ring[slot];
                ++lo;
            }
            ring[slot] = make_uint2(topkey, (uint32_t)topv | ((uint32_t)tops << 16));
        }
        ++k;
        topv = v; tops = start; topkey = key;
    }
    __device__ __forceinline__ void pop() {
        --k;
        if (k < 0) return;
        uint2 e;
        if (k >= lo) {
            e = ring[(k & (kRing - 1)) * 32];
        } else {                                 // ring empty: fetch from the spilled part
            e = spill[k];
            lo = k;
        }
        topkey = e.x;
        topv = (int)(e.y & 0xFFFFu);
        tops = (int)(e.y >> 16);
    }
};

// kFromG = false: g is derived from the band records of dt_col_band_kernel (product path).
// kFromG = true : g rows are given explicitly (u16, 0xFFFF = FLT_MAX), any content (row-pass parity tests).
template <bool kFromG>
__global__ void __launch_bounds__(kBandWarps * 32) dt_row_band_kernel(const uint2* __restrict__ info,
                                                                      const uint16_t* __restrict__ g, MapDims dm, int nbands,
                                                                      uint2* __restrict__ spill_all, int maxdepth,
                                                                      int32_t* __restrict__ row_k) {
    __shared__ __align__(16) uint2 ring_all[kBandWarps][kRing * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kBandWarps + warp;          // band id, plane-fastest (neighbouring warps: different planes)
    if (wg >= dm.D * nbands) return;
    const int d = wg % dm.D, b = wg / dm.D;
    const int row0 = b * 32;
    const int W = dm.W;
    const uint2* info_row = info + ((size_t)d * nbands + b) * dm.pitch;
    const uint16_t* g_row = kFromG ? g + ((size_t)d * dm.H + min(row0 + lane, dm.H - 1)) * dm.pitch : nullptr;
    const bool row_ok = row0 + lane < dm.H;

    RowStack st;
    st.ring = ring_all[warp] + lane;
    st.spill = spill_all + ((size_t)d * dm.H + min(row0 + lane, dm.H - 1)) * maxdepth;   // rows >= H never push
    st.k = -1; st.lo = 0; st.topkey = 0; st.topv = 0; st.tops = 0;
    const uint32_t mle = 0xFFFFFFFFu >> (31 - lane), mge = 0xFFFFFFFFu << lane;

    // ---- lower envelope over the columns that hold a finite g (imgproc.h:101-121) ----
    for (int x0 = 0; x0 < dm.pitch; x0 += 32) {
        uint2 e = make_uint2(0u, 0xFFFFFFFFu);
        unsigned todo;
        if (kFromG) {
            // lane = column here: which of the 32 columns hold a finite value in ANY row of the band
            unsigned any = 0;
            for (int r = 0; r < 32 && row0 + r < dm.H; ++r)
                if (g[((size_t)d * dm.H + row0 + r) * dm.pitch + x0 + lane] != kNone16) any = 1;
            todo = __ballot_sync(0xffffffffu, any && x0 + lane < W);
        } else {
            if (x0 + lane < W) e = info_row[x0 + lane];
            todo = __ballot_sync(0xffffffffu, e.x != 0u || e.y != 0xFFFFFFFFu);
        }
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int v = x0 + j;
            int gv;
            bool fin = row_ok;
            if (kFromG) {
                gv = g_row[v];
                fin = row_ok && gv != (int)kNone16;
            } else {
                const uint32_t M = __shfl_sync(0xffffffffu, e.x, j), ud = __shfl_sync(0xffffffffu, e.y, j);
                const uint32_t above = M & mle, below = M & mge;
                const int upd = above ? (lane - (31 - __clz(above))) : (lane + (int)(ud & 0xFFFFu));
                const int dnd = below ? (__ffs(below) - 1 - lane) : (31 - lane + (int)(ud >> 16));
                gv = min(upd, dnd);
            }
            if (fin) {
                const uint32_t key = (uint32_t)(gv * gv) + (uint32_t)(v * v);
                int start = 0;
                while (st.k >= 0) {
                    // parabola v beats the top strictly from pixel floor(N / Dn) + 1 on (left one keeps ties)
                    const int N = (int)key - (int)st.topkey;
                    const int Dn = 2 * (v - st.topv);
                    if (N < st.tops * Dn) {      // ... which is not after the top's first pixel: the top owns nothing
                        st.pop();
                        continue;
                    }
                    const float qf = __fdividef((float)N, (float)Dn);
                    if (qf >= 4096.f) { start = 0x7FFF; break; }   // far beyond the row (W <= 2897): never an owner
                    int t = (int)qf;
                    const int r = N - t * Dn;
                    if (r < 0) --t; else if (r >= Dn) ++t;
                    start = t + 1;
                    break;
                }
                if (start < W) st.push(v, start, key);
            }
        }
    }
    // ---- park the whole stack in the row's global array; the fill kernel walks it ----
    const int K = st.k + 1;
    if (K > 0) {
        for (int i = st.lo; i < st.k; ++i) st.spill[i] = st.ring[(i & (kRing - 1)) * 32];
        st.spill[st.k] = make_uint2(st.topkey, (uint32_t)st.topv | ((uint32_t)st.tops << 16));
    }
    if (row_ok) row_k[(size_t)d * dm.H + row0 + lane] = K;
}

// =============================================================================================
// fill (imgproc.h:122-128 incl. its in-place aliasing), one warp per row, lane = pixel of a 32-pixel chunk.
// The row's envelope (vertex v, first owned pixel s, f(v) + v^2; s strictly increasing) is read 32 entries at a time,
// one entry per lane.  The entries whose interval starts inside the chunk set one bit each in `marks`; the owner of
// pixel p is the popc(marks & bits <= p)-th of them (or the entry carried over from the left).  The value a vertex
// contributes is f(v) when it lies inside its own interval (s <= v) and the ALREADY WRITTEN out(v) when it lies left
// of it (s > v): out(v) comes from the shared-memory copy of the row (earlier chunk) or from another lane of the same
// chunk (resolved iteratively; the dependency always points to a lower entry index).
// =============================================================================================
constexpr int kFillWarps = 4;

__global__ void __launch_bounds__(kFillWarps * 32) dt_row_fill_kernel(const uint2* __restrict__ spill_all,
                                                                      const int32_t* __restrict__ row_k,
                                                                      float* __restrict__ planes, MapDims dm, int n_rows_total,
                                                                      int maxdepth, int win_lo, int win_w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kFillWarps + warp;           // row of the [D*H][pitch] stack of planes
    if (row >= n_rows_total) return;
    uint32_t* rowbuf = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * win_w - win_lo;   // indexed by absolute column
    const int W = dm.W;
    float* orow = planes + (size_t)row * dm.pitch;
    const int K = row_k[row];
    if (K <= 0) {                                             // no edge pixel in this plane: FLT_MAX stays (imgproc.h:174)
        for (int q = lane; q < W; q += 32) orow[q] = FLT_MAX;
        return;
    }
    const uint2* sp = spill_all + (size_t)row * maxdepth;
    const uint2 kSentinel = make_uint2(0u, 0xFFFFFFFFu);      // s = 0xFFFF: never starts inside a chunk
    uint2 be = lane < K ? sp[lane] : kSentinel;
    uint2 nbe = 32 + lane < K ? sp[32 + lane] : kSentinel;
    int e0 = 0;
    int bv = (int)(be.y & 0xFFFFu), bs = (int)(be.y >> 16);
    uint32_t bf = be.x - (uint32_t)(bv * bv);
    int carry_v = __shfl_sync(0xffffffffu, bv, 0);            // entry 0: s = 0 <= v
    uint32_t carry_b = __shfl_sync(0xffffffffu, bf, 0);
    const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);
    for (int q0 = 0; q0 < dm.pitch; q0 += 32) {
        const int q = q0 + lane;
        int ov = carry_v;
        uint32_t ob = carry_b;
        while (true) {
            const int rel = bs - q0;
            const bool inb = (unsigned)rel < 32u;
            const unsigned marks = __reduce_or_sync(0xffffffffu, inb ? (1u << rel) : 0u);
            if (marks) {
                const unsigned bal = __ballot_sync(0xffffffffu, inb);
                const int first = __ffs(bal) - 1;
                const bool chain = inb && bs > bv;
                bool ready = !chain;
                uint32_t base = bf;
                if (chain && bv < q0) {
                    base = rowbuf[bv];
                    ready = true;
                }
                const int c = __popc(marks & le_mask);
                const int src = (first + c - 1) & 31;
                unsigned pend = __ballot_sync(0xffffffffu, inb && !ready);
                int pv;
                uint32_t pb;
                while (true) {
                    const uint32_t sb = __shfl_sync(0xffffffffu, base, src);
                    const int sv = __shfl_sync(0xffffffffu, bv, src);
                    pv = c ? sv : ov;
                    pb = c ? sb : ob;
                    if (!pend) break;
                    const int dq = q - pv;
                    const uint32_t val = pb + (uint32_t)(dq * dq);
                    const bool pok = c == 0 || !((pend >> src) & 1u);       // this pixel's owner already has its base
                    const int pix = (bv - q0) & 31;
                    const uint32_t vv = __shfl_sync(0xffffffffu, val, pix);
                    const unsigned okb = __ballot_sync(0xffffffffu, pok);
                    if (inb && !ready && ((okb >> pix) & 1u)) {
                        base = vv;
                        ready = true;
                    }
                    pend = __ballot_sync(0xffffffffu, inb && !ready);
                }
                ov = pv;
                ob = pb;
                const int lastl = 31 - __clz(bal);
                carry_v = __shfl_sync(0xffffffffu, bv, lastl);
                carry_b = __shfl_sync(0xffffffffu, base, lastl);
            }
            const int last_s = __shfl_sync(0xffffffffu, bs, 31);
            if (last_s < q0 + 32 && e0 + 32 < K) {            // the chunk continues in the next 32 entries
                e0 += 32;
                be = nbe;
                nbe = e0 + 32 + lane < K ? sp[e0 + 32 + lane] : kSentinel;
                bv = (int)(be.y & 0xFFFFu);
                bs = (int)(be.y >> 16);
                bf = be.x - (uint32_t)(bv * bv);
                continue;
            }
            break;
        }
        const int dq = q - ov;
        const uint32_t val = ob + (uint32_t)(dq * dq);
        if (q >= win_lo && q < win_lo + win_w) rowbuf[q] = val;
        if (q < W) orow[q] = (float)val;                      // < 2^24: exact
        __syncwarp();
    }
}

static inline unsigned cdiv_u(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

int dt_band_count(const MapDims& dm) { return (dm.H + 31) / 32; }
size_t dt_band_info_bytes(const MapDims& dm) { return (size_t)dm.D * dt_band_count(dm) * dm.pitch * sizeof(uint2); }
static size_t row_k_bytes(const MapDims& dm) { return ((size_t)dm.D * dm.H * sizeof(int32_t) + 255) / 256 * 256; }
// workspace of the row call: per-row envelope length + per-row envelope array of maxdepth entries
size_t dt_band_spill_bytes(const MapDims& dm, int maxdepth) {
    return row_k_bytes(dm) + (size_t)dm.D * dm.H * (size_t)maxdepth * sizeof(uint2);
}

void launch_dt_col_band(const uint32_t* d_mask, const MapDims& dm, void* d_info, cudaStream_t s) {
    const int nbands = dt_band_count(dm);
    const size_t smem = (size_t)nbands * 64 * 6;
    dim3 grid((dm.wwords + 1) / 2, dm.D);
    dt_col_band_kernel<<<grid, 256, smem, s>>>(d_mask, dm, reinterpret_cast<uint2*>(d_info), nbands);
}

// [win_lo, win_hi]: columns that can hold an edge pixel (envelope vertices only exist there)
struct RowWs { int32_t* row_k; uint2* spill; int win_lo, maxdepth; };
static RowWs row_ws(const MapDims& dm, void* d_ws, int win_lo, int win_hi) {
    win_lo = win_lo < 0 ? 0 : win_lo;
    win_hi = win_hi >= dm.W ? dm.W - 1 : win_hi;
    if (win_hi < win_lo) { win_lo = 0; win_hi = dm.W - 1; }
    return RowWs{reinterpret_cast<int32_t*>(d_ws), reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(d_ws) + row_k_bytes(dm)),
                 win_lo, win_hi - win_lo + 1};
}

void launch_dt_row_envelope(const void* d_info, const uint16_t* d_g, const MapDims& dm, void* d_ws, int win_lo, int win_hi,
                            cudaStream_t s) {
    const RowWs ws = row_ws(dm, d_ws, win_lo, win_hi);
    const int nbands = dt_band_count(dm);
    const unsigned grid = cdiv_u((size_t)dm.D * nbands, kBandWarps);
    if (d_g)
        dt_row_band_kernel<true><<<grid, kBandWarps * 32, 0, s>>>(nullptr, d_g, dm, nbands, ws.spill, ws.maxdepth, ws.row_k);
    else
        dt_row_band_kernel<false><<<grid, kBandWarps * 32, 0, s>>>(reinterpret_cast<const uint2*>(d_info), nullptr, dm, nbands, ws.spill,
                                                                  ws.maxdepth, ws.row_k);
}

void launch_dt_row_fill(float* d_planes, const MapDims& dm, void* d_ws, int win_lo, int win_hi, cudaStream_t s) {
    const RowWs ws = row_ws(dm, d_ws, win_lo, win_hi);
    const int rows = dm.D * dm.H;
    const size_t smem = (size_t)kFillWarps * ws.maxdepth * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(dt_row_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    dt_row_fill_kernel<<<cdiv_u(rows, kFillWarps), kFillWarps * 32, smem, s>>>(ws.spill, ws.row_k, d_planes, dm, rows, ws.maxdepth,
                                                                              ws.win_lo, ws.maxdepth);
}

}   // namespace fdcm
