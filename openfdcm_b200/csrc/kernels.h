// kernels.h — launcher declarations shared by the C-ABI layer (fdcm_api.cu)
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the driver entry point is resolved at run time)

#include "common.cuh"
#include "../../include/fdcm_b200.h"

namespace fdcm {

// ---- dt3_kernels.cu ----
void launch_raster(const float* d_lines, const int32_t* d_bins, int n_lines, const MapDims& dm, uint32_t* d_mask, cudaStream_t s);
// literal Felzenszwalb pass; src_kind: 0 = the plane itself, 1 = the 1-bit edge mask (column pass only), 2 = explicit u16 distances
void launch_dt_pass_literal(int src_kind, bool along_rows, const void* d_src, float* d_planes, const MapDims& dm, void* d_stack,
                            cudaStream_t s);
void launch_transpose_square(float* d_planes, const MapDims& dm, cudaStream_t s);
void launch_sqrt(float* d_planes, const MapDims& dm, cudaStream_t s);
void launch_propagate(float* d_planes, const MapDims& dm, const PropParams& pp, bool sqrt_first, cudaStream_t s);

// ---- integral_tma.cu: lineIntegral as one persistent TMA-fed kernel ----
struct IntegralPlanDev {
    const int4* items4;          // (plane, first chain of the strip, first tile, end tile: where the strip meets the image), heaviest first
    int n_items;
    int* counter;                // work counter, zero before the launch
    CUtensorMap map_y, map_x;    // the planes with the y-major / x-major box shapes
};
size_t integral_tma_smem_bytes();
int integral_strip_chains();
int integral_tile_steps();
bool integral_tma_encode(const void* planes, const MapDims& dm, CUtensorMap* map_y, CUtensorMap* map_x);
void launch_integral_tma(float* d_planes, const MapDims& dm, const IntegralParams& ip, const IntegralPlanDev& plan, int n_sms,
                         cudaStream_t s);

// ---- dt_band_kernels.cu (exact regime, lane-per-row formulation) ----
int dt_band_count(const MapDims& dm);
size_t dt_band_info_bytes(const MapDims& dm);
size_t dt_band_spill_bytes(const MapDims& dm, int maxdepth);
void launch_dt_col_band(const uint32_t* d_mask, const MapDims& dm, void* d_info, cudaStream_t s);
// d_info (band records) or d_g (explicit u16 rows, tests) feeds the envelope build; exactly one of them is non-null
// [row_lo, row_hi]: rows that can hold edge pixels (only used to order the bands in the grid)
// which: 0 = every band, 1 = only the bands overlapping the scene rows [row_lo, row_hi], 2 = only the other (far) bands
void launch_dt_row_envelope(const void* d_info, const uint16_t* d_g, const MapDims& dm, void* d_ws, int win_lo, int win_hi, int row_lo,
                            int row_hi, int which, cudaStream_t s);
void dt_band_scene_rows(const MapDims& dm, int row_lo, int row_hi, int* y0, int* y1);
void launch_dt_row_fill(float* d_planes, const MapDims& dm, void* d_ws, int win_lo, int win_hi, cudaStream_t s);
// fill fused with propagateOrientation (and the L2 sqrt): the distance-transform planes never reach HBM
bool dt_fill_propagate_supported(const MapDims& dm);
// image rows [ya0, ya1) and [yb0, yb1); the resolve pass rewrites the envelope arrays in place and must precede the fused fill
void launch_dt_resolve(const MapDims& dm, void* d_ws, int win_lo, int win_hi, int ya0, int ya1, int yb0, int yb1, cudaStream_t s);
void launch_dt_fill_propagate(float* d_planes, const MapDims& dm, void* d_ws, int win_lo, int win_hi, const PropParams& pp,
                              bool sqrt_first, int ya0, int ya1, int yb0, int yb1, cudaStream_t s);

// L1: row call straight from the band records (exact at any size), stand-alone or fused with propagateOrientation
void launch_dt_row_l1_band(const void* d_info, float* d_planes, const MapDims& dm, cudaStream_t s);
void launch_dt_l1_propagate(const void* d_info, float* d_planes, const MapDims& dm, const PropParams& pp, cudaStream_t s);
size_t dt_band_smem_bytes(const MapDims& dm);
// parity hook of the fused fill's square root (integers 0 .. 2^24 and FLT_MAX against the IEEE sqrtf)
cudaError_t run_sqrt_check(unsigned long long* h_bad, uint32_t* h_first, cudaStream_t s);

// ---- search_kernels.cu ----
struct MapView {                 // read-only view of a built feature map
    const float* planes;
    MapDims dm;
    float shift_x, shift_y;      // sceneTranslation (dt3cpu.h:56)
};

struct TemplatesView {
    const float4* lines;         // all template lines, CSR by offsets
    const int32_t* offsets;      // n_tmpl + 1
    const int32_t* argsort;      // per template: line indices by descending length (std::sort order)
    const float* line_len;       // per line length
    const float* denom;          // per template penalty denominator (or nullptr)
    int32_t n_tmpl;
    int32_t max_lines;           // longest template
};

struct SceneView {
    const float4* lines;         // original (un-shifted) scene lines
    const float* sorted_len;     // lengths, descending (std::sort order of the reference comparator)
    const int32_t* sorted_idx;   // scene line index of each sorted entry
    int32_t n;
};

struct SearchLaunch {
    int32_t max_tmpl_lines, max_scene_lines, batch;
    int32_t tmpl_idx_base;
    const int64_t* hyp_off;      // n_tmpl + 1 prefix of hypothesis counts
    int64_t n_hyp;
    const int32_t* perm;         // processing order: slot i handles hypothesis perm[i] (nullptr: identity)
    const int4* hyp_ready;       // hypotheses already decoded by the ordering pass (tmpl + base, tmpl line, scene line, rev)
    const float2* direct_align;  // optimize() entry point: hypothesis h = template h as given (identity transform)
                                 // with alignment vector direct_align[h]; nullptr for the fused search
};

struct SearchOutputs {
    fdcm_match* rec;             // n_hyp records
    uint8_t* valid;              // n_hyp flags (0 = nullopt)
    int4* hyp;                   // n_hyp (tmpl, tmpl_line, scene_line, rev)
    unsigned long long* counters;// [0] evaluations, [1] lookups, [2] valid
};

// spatial ordering of the hypotheses (locality of the map gathers): keys + radix sort -> permutation
// Cell of the spatial hypothesis order (search_key_kernel).  The hypotheses are processed cell row by cell row, x fastest:
// the map rows a cell row reaches (its own height + twice the template radius, all D planes) slide through the L2 once per
// cell row, so tall cells re-read the map fewer times, as long as the window (rows reached x columns of the cells in
// flight) still fits the L2.  Measured on config 3 (1080p, D = 30, templates reaching +-270 px), search kernel alone, cell
// height 128 -> 1.174 ms, 256 -> 1.149, 336 -> 1.157, 352 -> 1.053, 368 -> 1.076, 384 -> 1.071, 400 -> 1.073, 416 -> 1.148,
// 448 -> 1.102, 480 -> 1.125, 512 -> 1.145, 640 -> 1.051, 720 -> 1.067, one row of 32-px columns -> 1.090: tall beats short,
// with +-4 % of scene-dependent scatter on top (a height derived from the L2 size, the depth and the template diagonal was
// tried and landed on 480: the scatter is larger than what such a model resolves, so the height is a constant).
constexpr int kSearchCellW = 128, kSearchCellH = 384;
size_t search_order_temp_bytes(int64_t n_hyp);
void launch_search_order(const TemplatesView& tv, const SceneView& sv, const SearchLaunch& sl, uint32_t* d_keys, uint32_t* d_keys_out,
                         int32_t* d_idx, int32_t* d_perm, void* d_temp, size_t temp_bytes, float minx, float miny, int cells_x,
                         int key_bits, int4* d_hyp, cudaStream_t s);

void launch_search(const MapView& map, const SlopeTableDev& table, const TemplatesView& tv, const SceneView& sv,
                   const SearchLaunch& sl, const SearchOutputs& out, cudaStream_t s);

// top-K smallest (score, index) among valid records; returns via d_out (k records) and d_n_out
void launch_topk(const fdcm_match* d_rec, const uint8_t* d_valid, int64_t n, int k, float* d_ws_score, int64_t* d_ws_idx,
                 int ws_blocks, fdcm_match* d_out, int* d_n_out, cudaStream_t s);
int topk_ws_blocks(int64_t n);
// multi-GPU: merge of the all-gathered per-rank top-k lists; k invalid records for an empty shard
void launch_topk_merge(const fdcm_match* d_gathered, int n_cand, int k, fdcm_match* d_out, int* d_n_out, cudaStream_t s);
void launch_topk_invalid(fdcm_match* d_out, int k, int* d_n_out, cudaStream_t s);

void launch_evaluate(const MapView& map, const SlopeTableDev& table, const float4* d_lines, const int32_t* d_toff,
                     const float2* d_transl, const int32_t* d_troff, const int32_t* d_owner, int64_t n_scores,
                     float* d_scores, cudaStream_t s);
void launch_classify(const SlopeTableDev& table, const float4* d_lines, int n, int32_t* d_bins, cudaStream_t s);

}   // namespace fdcm
