// Microbenchmark: throughput of red.shared.add.u32 on sm_100a for the access patterns the pileup kernel uses.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void red(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
template <int MODE>
__global__ void __launch_bounds__(256, 4) k(uint32_t *out, int iters, uint32_t seed) {
    __shared__ uint32_t s[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) s[i] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
    const int lane = threadIdx.x & 31;
    uint32_t x = seed + threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx;
        if (MODE == 0) idx = (x >> 8) & 8191u;                               // random address, all lanes
        else if (MODE == 1) idx = ((i * 37 + lane) & 8191u);                 // distinct banks, all lanes
        else if (MODE == 2) idx = ((x >> 8) & 8191u);                        // random, ~5/32 lanes active (predicated by branch)
        else if (MODE == 3) idx = ((i * 5 + lane * 5) & 8191u);              // 5 consecutive words per lane (SWAR pattern), all lanes
        else idx = (i & 8191u);                                              // same address all lanes
        if (MODE == 2) { if (((x >> 27) & 31u) < 5u) red(base + 4 * idx, 1u); }
        else if (MODE == 3) { red(base + 4 * idx, x & 1u); red(base + 4 * ((idx + 1) & 8191u), 0u); red(base + 4 * ((idx + 2) & 8191u), 0u);
                              red(base + 4 * ((idx + 3) & 8191u), 0u); red(base + 4 * ((idx + 4) & 8191u), 0u); }
        else red(base + 4 * idx, 1u);
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (uint32_t)(t1 - t0) + s[1] * 0;
}
template <int MODE> void run(const char *name, int atoms_per_iter) {
    uint32_t *d; cudaMalloc(&d, 148 * 4 * 4);
    const int iters = 20000;
    k<MODE><<<148 * 4, 256>>>(d, iters, 1); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<148 * 4, 256>>>(d, iters, 2); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    uint32_t h[592]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 592; i++) cyc += h[i]; cyc /= 592;
    // per SM: 32 warps each issuing iters*atoms_per_iter warp-level ATOMS
    const double warp_atoms_per_sm = 32.0 * iters * atoms_per_iter;
    printf("%-44s %8.3f ms  %9.0f cyc/CTA-loop  => %.2f cycles per warp-ATOMS per SM (%.2f ATOMS/clk/SM)\n", name, ms, cyc, cyc / warp_atoms_per_sm * 1.0,
           warp_atoms_per_sm / cyc);
    cudaFree(d);
}
int main() {
    run<0>("random addr, 32 lanes", 1);
    run<1>("distinct banks, 32 lanes", 1);
    run<2>("random addr, ~5/32 lanes (branch)", 1);
    run<3>("5 consecutive words/lane, 32 lanes, zeros", 5);
    run<4>("same address, 32 lanes", 1);
    return 0;
}
