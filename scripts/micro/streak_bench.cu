// Micro-benchmark for the search gather pattern: 4-byte lookups along 10-pixel streaks (BatchOptimize candidates) in a
// [30][2880][2880] fp32 map.  mode 0: lane = streak, 10 sequential gathers (today's search8 pattern);
// mode 1: 3 streaks x 10 lanes per instruction, row-major map; mode 2: same on an 8x4-tiled map.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int S = 2880, D = 30;
__device__ __forceinline__ size_t addr_rm(int d, int x, int y) { return ((size_t)d * S + y) * S + x; }
__device__ __forceinline__ size_t addr_t84(int d, int x, int y) {
    return (((size_t)d * (S / 4) + (y >> 2)) * (S / 8) + (x >> 3)) * 32 + ((y & 3) << 3) + (x & 7);
}
__global__ void gather(const float* __restrict__ p, int mode, int iters, float* out) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f;
    unsigned long long s = warp * 0x9E3779B97F4A7C15ull + 12345;
    for (int it = 0; it < iters; ++it) {
        // a "hypothesis neighbourhood": all streaks of this iteration start within a 200x200 window (template extent)
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const int wx = 480 + (int)((s >> 33) % 1700), wy = 900 + (int)((s >> 13) % 860);
        const bool xmaj = (s >> 7) & 1;
        const float slope = (float)((s >> 40) & 1023) / 1023.f;
        if (mode == 0) {
            unsigned long long t = s ^ (0x9E3779B1ull * (lane + 1));
            t = t * 6364136223846793005ull + 1442695040888963407ull;
            const int d = (t >> 50) % D, bx = wx + (int)((t >> 20) % 200), by = wy + (int)((t >> 36) % 200);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const int x = xmaj ? bx + j : bx + (int)(j * slope), y = xmaj ? by + (int)(j * slope) : by + j;
                acc += __ldg(p + addr_rm(d, x, y));
            }
        } else {
            const int j = lane % 10, sl = lane / 10;
            for (int k = 0; k < 11; ++k) {      // 32 streaks -> 11 instructions of 3 streaks
                unsigned long long t = s ^ (0x9E3779B1ull * (k * 3 + sl + 1));
                t = t * 6364136223846793005ull + 1442695040888963407ull;
                const int d = (t >> 50) % D, bx = wx + (int)((t >> 20) % 200), by = wy + (int)((t >> 36) % 200);
                const int x = xmaj ? bx + j : bx + (int)(j * slope), y = xmaj ? by + (int)(j * slope) : by + j;
                if (sl < 3) acc += __ldg(p + (mode == 1 ? addr_rm(d, x, y) : addr_t84(d, x, y)));
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    const size_t bytes = (size_t)D * S * S * 4;
    float* p; float* out;
    cudaMalloc(&p, bytes); cudaMalloc(&out, 4); cudaMemset(p, 0, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode) {
        const int blocks = 148 * 8, threads = 128, iters = 300;
        gather<<<blocks, threads>>>(p, mode, 10, out);
        cudaEventRecord(a);
        gather<<<blocks, threads>>>(p, mode, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double lookups = (double)blocks * threads / 32 * iters * (mode == 0 ? 320.0 : 330.0);
        printf("mode %d: %.3f ms, %.1f G lookups/s\n", mode, ms, lookups / ms / 1e6);
    }
    return 0;
}
