// Micro-benchmark: achievable NVLink bandwidth between two B200s of one box for the access patterns the sharded
// construction uses (one process, plain peer access).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/peer_bw.cu -o tools/peer_bw
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;

__global__ void st8(u64* dst, const u64* src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void st16(ulonglong2* dst, const ulonglong2* src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// every warp writes `run` consecutive u64 at a pseudo-random run-aligned place (the digit runs of a radix pass)
__global__ void st8_runs(u64* dst, const u64* src, size_t n, int run) {
    const size_t nruns = n / run;
    const int lane = threadIdx.x & 31;
    for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nruns; w += ((size_t)gridDim.x * blockDim.x) >> 5) {
        const size_t r = (w * 2654435761ull) % nruns;
        for (int j = lane; j < run; j += 32) dst[r * run + j] = src[w * run + j];
    }
}
__global__ void ld16(ulonglong2* dst, const ulonglong2* src, size_t n) {  // pull: remote loads, local stores
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// random 8-byte gathers from the peer (bulk get of ISA entries)
__global__ void gather8(u64* dst, const u64* src, size_t n, size_t m) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[(i * 0x9E3779B97F4A7C15ull) % n];
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t bytes = (size_t)4 << 30, n = bytes / 8;
    u64 *a0, *b0, *a1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, bytes)); CK(cudaMemset(a1, 1, bytes)); CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, bytes)); CK(cudaMalloc(&b0, bytes)); CK(cudaMemset(a0, 2, bytes)); CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 8, thr = 512;
    auto report = [&](const char* name, double b) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); printf("%-46s %8.3f ms  %8.1f GB/s\n", name, ms, b / ms / 1e6); };
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0)); st8<<<grid, thr>>>(b0, a0, n); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("local  8-byte stores (copy)", bytes);
        CK(cudaEventRecord(e0)); st8<<<grid, thr>>>(a1, a0, n); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("peer   8-byte stores, coalesced", bytes);
        CK(cudaEventRecord(e0)); st16<<<grid, thr>>>((ulonglong2*)a1, (const ulonglong2*)a0, n / 2); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("peer  16-byte stores, coalesced", bytes);
        for (int run : {24, 32, 256, 4096}) {
            char nm[64]; snprintf(nm, 64, "peer   8-byte stores, runs of %d words", run);
            CK(cudaEventRecord(e0)); st8_runs<<<grid, thr>>>(a1, a0, n, run); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report(nm, (double)(n / run) * run * 8);
        }
        CK(cudaEventRecord(e0)); ld16<<<grid, thr>>>((ulonglong2*)b0, (const ulonglong2*)a1, n / 2); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("peer  16-byte loads (pull), coalesced", bytes);
        CK(cudaEventRecord(e0)); CK(cudaMemcpyPeerAsync(a1, 1, a0, 0, bytes)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("cudaMemcpyPeerAsync (copy engine)", bytes);
        const size_t m = (size_t)1 << 26;
        CK(cudaEventRecord(e0)); gather8<<<grid, thr>>>(b0, a1, n, m); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("peer random 8-byte gathers (2^26)", (double)m * 8);
        CK(cudaEventRecord(e0)); gather8<<<grid, thr>>>(b0, a0, n, m); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); report("local random 8-byte gathers (2^26)", (double)m * 8);
    }
    // both directions at once: GPU1 stores into GPU0 while GPU0 stores into GPU1
    u64* b1; CK(cudaSetDevice(1)); CK(cudaMalloc(&b1, bytes)); CK(cudaMemset(b1, 3, bytes));
    cudaStream_t s1; CK(cudaStreamCreate(&s1));
    CK(cudaSetDevice(0));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    st16<<<grid, thr>>>((ulonglong2*)a1, (const ulonglong2*)a0, n / 2);
    CK(cudaSetDevice(1)); st16<<<grid, thr, 0, s1>>>((ulonglong2*)b0, (const ulonglong2*)b1, n / 2); CK(cudaSetDevice(0));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0));
    report("peer 16-byte stores, both directions (GPU0 side)", bytes);
    return 0;
}
