// Stand-alone timing of ONE radix digit pass (u32 keys, u32 values [, aux byte]: histogram + scan + scatter) with per-phase cycle
// accounting of the scatter kernel.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DPSAC_PHASE_PROFILE -o tools/bench_pass tools/bench_pass.cu
#include <cstdio>
#include <vector>
#include "../psac_b200/csrc/radix_sort.cuh"
using namespace psacb200;
void psacb200::set_last_error(const std::string&) {}

__global__ void fill(u32* k, u32* v, u8* a, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 x = i * 0x9E3779B97F4A7C15ull + 12345;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        k[i] = (u32)x; v[i] = (u32)i; a[i] = (u8)(x >> 40);
    }
}

template <bool AUX, int THREADS, int ITEMS, int MINB>
void run(size_t n, int reps) {
    u32 *k[2], *v[2]; u8* a[2];
    for (int b = 0; b < 2; ++b) { cudaMalloc(&k[b], n * 4); cudaMalloc(&v[b], n * 4); cudaMalloc(&a[b], n); }
    fill<<<1184, 256>>>(k[0], v[0], a[0], n);
    RadixWorkspace ws;
    cudaMalloc(&ws.gbase, RADIX * 8);
    ws.tiles_bytes = RadixWorkspace::tiles_bytes_for(n); cudaMalloc(&ws.tiles, ws.tiles_bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    unsigned long long zero[16] = {0};
    for (int r = 0; r < reps; ++r) {
        cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero));
        ArraySrc<u32, u32> src{k[0], v[0], AUX ? a[0] : nullptr, 8, 255u, 0u};
        cudaEventRecord(e0);
        launch_pass_cfg<ArraySrc<u32, u32>, u32, AUX, THREADS, ITEMS, MINB>(ws, src, k[1], v[1], AUX ? a[1] : nullptr, n, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    unsigned long long ph[16]; cudaMemcpyFromSymbol(ph, g_phase_cycles, sizeof(ph));
    const double tiles = (double)((n + THREADS * ITEMS - 1) / (THREADS * ITEMS));
    printf("n=%zu aux=%d threads=%d items=%d minb=%d: %.3f ms  (%.1f GB/s algorithmic)  err=%s\n", n, (int)AUX, THREADS, ITEMS, MINB, best, n * (AUX ? 18.0 : 16.0) / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    const char* names[7] = {"key loads + count", "barrier 1", "digit scan", "rank + scatter", "offsets", "barrier 3", "write-out"};
    double tot = 0; for (int i = 0; i < 7; ++i) tot += ph[i];
    for (int i = 0; i < 7; ++i) printf("   %-18s %9.0f cycles/tile  %5.1f%%\n", names[i], ph[i] / tiles, 100.0 * ph[i] / tot);
    printf("   total %.0f cycles per tile (thread 0), tiles=%.0f\n", tot / tiles, tiles);
    for (int b = 0; b < 2; ++b) { cudaFree(k[b]); cudaFree(v[b]); cudaFree(a[b]); }
    cudaFree(ws.gbase); cudaFree(ws.tiles);
}

int main() {
    const size_t n = (size_t)1 << 28;
    run<true, 384, 12, 3>(n, 3);
    run<true, 384, 16, 2>(n, 3);
    run<true, 512, 12, 2>(n, 3);
    run<true, 512, 16, 2>(n, 3);
    run<true, 384, 20, 2>(n, 3);
    run<true, 640, 12, 1>(n, 3);
    run<true, 768, 12, 1>(n, 3);
    run<true, 1024, 12, 1>(n, 3);
    run<false, 384, 12, 3>(n, 3);
    run<false, 384, 16, 2>(n, 3);
    run<false, 512, 12, 2>(n, 3);
    run<false, 512, 16, 2>(n, 3);
    return 0;
}
