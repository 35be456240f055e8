// Micro-benchmark: cost of a warp-level 4-byte gather as a function of how many distinct 128-byte lines / 32-byte
// sectors the 32 lanes touch (decides whether tiling the DT3 map + lanes-as-candidates can speed up search8_kernel).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void gather(const float* __restrict__ p, size_t n_lines, int mode, int iters, float* out) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long s = warp * 0x9E3779B97F4A7C15ull + 12345;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            size_t line;
            int off;   // float offset inside the 128-byte line
            if (mode == 0) {            // 32 lanes -> 32 distinct random lines
                line = ((s >> 20) + (unsigned long long)lane * 0x9E3779B1ull) % n_lines; off = (s >> 8) & 31;
            } else if (mode == 1) {     // 8 random lines x 4 lanes, each lane a different sector of its line
                line = ((s >> 20) + (unsigned long long)(lane >> 2) * 0x9E3779B1ull) % n_lines; off = (lane & 3) * 8 + ((s >> 8) & 7);
            } else if (mode == 2) {     // 8 random lines x 4 lanes, same sector
                line = ((s >> 20) + (unsigned long long)(lane >> 2) * 0x9E3779B1ull) % n_lines; off = ((s >> 8) & 3) * 8 + (lane & 3);
            } else {                    // 16 random lines x 2 lanes, same sector
                line = ((s >> 20) + (unsigned long long)(lane >> 1) * 0x9E3779B1ull) % n_lines; off = ((s >> 8) & 3) * 8 + (lane & 1);
            }
            v[k] = __ldg(p + line * 32 + off);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    const size_t bytes = 256ull << 20;   // 256 MB region (like the scene part of the map), partly L2 resident
    float* p; float* out;
    cudaMalloc(&p, bytes); cudaMalloc(&out, 4); cudaMemset(p, 0, bytes);
    const size_t n_lines = bytes / 128;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 4; ++mode) {
        const int blocks = 148 * 8, threads = 128, iters = 400;
        gather<<<blocks, threads>>>(p, n_lines, mode, 10, out);
        cudaEventRecord(a);
        gather<<<blocks, threads>>>(p, n_lines, mode, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double ldg = (double)blocks * threads / 32 * iters * 8;
        printf("mode %d: %.3f ms, %.1f M warp-gathers, %.2f ns per warp-gather per SM-slot, %.1f G lane-lookups/s\n", mode, ms, ldg / 1e6,
               ms * 1e6 / (ldg / 148), ldg * 32 / ms / 1e6);
    }
    return 0;
}
