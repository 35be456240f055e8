// rmw_bench.cu — which in-place read-modify-write access patterns over the [30][2880][2880] fp32 planes reach HBM speed?
// (design input for the lineIntegral kernel: strips of chains walking along y or x, tile by tile)
//   pattern: a CTA (256 threads) owns a strip and visits its tiles in order; a tile is R rows x C floats, read as
//   float4 (threads along the columns first), +1, written back; P tiles are kept in flight in registers.
//   ADV_X = false: strip = C columns, tiles advance down the rows (y-major chains)
//   ADV_X = true : strip = R rows, tiles advance along the columns (x-major chains)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rmw_bench rmw_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int W = 2880, H = 2880, D = 30, PITCH = 2880;

template <int R, int C, int P, bool ADV_X>
__global__ void __launch_bounds__(256) rmw_kernel(float* __restrict__ base, int n_strips, int n_items, int order) {
    constexpr int C4 = C / 4;
    constexpr int UNITS = R * C4;
    constexpr int UPT = (UNITS + 255) / 256;
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        // order 0: strips of one plane adjacent in the grid; 1: planes adjacent (same strip of every plane together)
        const int d = order == 0 ? item / n_strips : item % D;
        const int s = order == 0 ? item % n_strips : item / D;
        float* P0 = base + (size_t)d * H * PITCH;
        const int n_tiles = ADV_X ? W / C : H / R;
        auto tile_ptr = [&](int t, int u) -> float4* {
            const int r = u / C4, c4 = u % C4;
            const int y = ADV_X ? s * R + r : t * R + r;
            const int x = ADV_X ? t * C + 4 * c4 : s * C + 4 * c4;
            return reinterpret_cast<float4*>(P0 + (size_t)y * PITCH + x);
        };
        float4 buf[P][UPT];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                const int u = tid + k * 256;
                if (u < UNITS && p < n_tiles) buf[p][k] = *tile_ptr(p, u);
            }
        for (int t0 = 0; t0 < n_tiles; t0 += P) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int t = t0 + p;
                if (t < n_tiles) {
#pragma unroll
                    for (int k = 0; k < UPT; ++k) {
                        const int u = tid + k * 256;
                        if (u < UNITS) {
                            float4 v = buf[p][k];
                            v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
                            *tile_ptr(t, u) = v;
                            if (t + P < n_tiles) buf[p][k] = *tile_ptr(t + P, u);
                        }
                    }
                }
            }
        }
    }
}

// ceiling: contiguous in-place RMW, grid-stride float4
__global__ void __launch_bounds__(256) linear_kernel(float4* __restrict__ p, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256 * 4) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + (size_t)k * gridDim.x * 256 < n4) v[k] = p[i + (size_t)k * gridDim.x * 256];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + (size_t)k * gridDim.x * 256 < n4) {
            v[k].x += 1.f; v[k].y += 1.f; v[k].z += 1.f; v[k].w += 1.f;
            p[i + (size_t)k * gridDim.x * 256] = v[k];
        }
    }
}

template <int R, int C, int P, bool ADV_X>
static void run(float* d, int ctas_per_sm, int order) {
    const int n_strips = ADV_X ? H / R : W / C;
    const int n_items = n_strips * D;
    int grid = 148 * ctas_per_sm;
    if (grid > n_items) grid = n_items;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rmw_kernel<R, C, P, ADV_X><<<grid, 256>>>(d, n_strips, n_items, order);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) rmw_kernel<R, C, P, ADV_X><<<grid, 256>>>(d, n_strips, n_items, order);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double cover = (double)n_strips * (ADV_X ? R : C) * (ADV_X ? (W / C) * C : (H / R) * R) * D * 8.0;
    printf("%s R=%3d C=%4d P=%d ctas/sm=%d order=%d  items=%5d  %.3f ms  %.0f GB/s  err=%s\n", ADV_X ? "X" : "Y", R, C, P, ctas_per_sm, order,
           n_items, ms, cover / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* d;
    const size_t n = (size_t)D * H * PITCH;
    cudaMalloc(&d, n * 4);
    cudaMemset(d, 0, n * 4);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        linear_kernel<<<148 * 8, 256>>>((float4*)d, n / 4);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) linear_kernel<<<148 * 8, 256>>>((float4*)d, n / 4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("linear in-place RMW: %.3f ms  %.0f GB/s\n", ms, n * 8.0 / ms * 1e-6);
    }
    // y-major: strip width sweep (tile = R rows x C columns, advancing in y)
    for (int order = 0; order < 2; ++order) {
        run<32, 128, 2, false>(d, 4, order);
        run<32, 128, 4, false>(d, 4, order);
        run<16, 256, 4, false>(d, 4, order);
        run<8, 576, 4, false>(d, 4, order);
        run<8, 576, 4, false>(d, 2, order);
        run<4, 960, 4, false>(d, 4, order);
        run<4, 1440, 4, false>(d, 2, order);
        run<2, 2880, 4, false>(d, 2, order);
        run<2, 2880, 8, false>(d, 2, order);
    }
    // x-major: strip = R rows, tile width sweep (advancing in x)
    for (int order = 0; order < 2; ++order) {
        run<160, 32, 2, true>(d, 4, order);
        run<160, 32, 4, true>(d, 4, order);
        run<96, 64, 4, true>(d, 4, order);
        run<96, 64, 2, true>(d, 4, order);
        run<64, 96, 4, true>(d, 4, order);
        run<32, 192, 4, true>(d, 4, order);
        run<32, 192, 4, true>(d, 2, order);
        run<16, 360, 4, true>(d, 4, order);
    }
    cudaFree(d);
    return 0;
}
