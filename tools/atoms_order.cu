// Hardware-behaviour probe (sm_100a): when several lanes of ONE warp instruction do atomicAdd(+1) with return on the
// same shared-memory word, in which order are the lanes serialised?  If it is ascending lane order, the returned value
// is a STABLE rank (running count + number of lower lanes with the same digit) and a radix pass can rank a key with a
// single ATOMS instead of the atomicOr + LDS + STS sequence.  Counts mismatches against the ballot-computed stable rank.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atoms_order tools/atoms_order.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int NDIG>
__global__ void probe(unsigned long long* mism_asc, unsigned long long* mism_desc, unsigned long long* total, unsigned seed, int iters) {
    __shared__ unsigned tab[12][256];
    __shared__ unsigned ref[12][256];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 12 * 256; i += blockDim.x) { (&tab[0][0])[i] = 0; (&ref[0][0])[i] = 0; }
    __syncthreads();
    unsigned x = seed * 2654435761u + (blockIdx.x * blockDim.x + threadIdx.x) * 40503u + 12345u;
    unsigned long long bad_a = 0, bad_d = 0, tot = 0;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const unsigned d = (x >> 20) % NDIG;
        const bool active = ((x >> 8) & 15) != 0 || (it & 1);  // some rounds have inactive lanes (ragged tiles)
        unsigned got = 0xffffffffu;
        if (active) got = atomicAdd(&tab[warp][d], 1u);
        __syncwarp();
        // stable rank by ballots
        unsigned peers = __ballot_sync(0xffffffffu, active);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1);
            peers &= ((d >> b) & 1) ? m : ~m;
        }
        if (active) {
            const unsigned before = ref[warp][d];
            __syncwarp(peers);
            const unsigned asc = before + __popc(peers & lanemask_lt());
            const unsigned desc = before + __popc(peers & ~lanemask_lt() & ~(1u << lane));
            if ((peers & lanemask_lt()) == 0) ref[warp][d] = before + __popc(peers);
            bad_a += got != asc;
            bad_d += got != desc;
            tot += 1;
        }
        __syncwarp();
    }
    atomicAdd(mism_asc, bad_a);
    atomicAdd(mism_desc, bad_d);
    atomicAdd(total, tot);
}

template <int NDIG>
void run(int threads, int grid, int iters) {
    unsigned long long *d, h[3];
    cudaMalloc(&d, 24);
    cudaMemset(d, 0, 24);
    probe<NDIG><<<grid, threads>>>(d, d + 1, d + 2, 7u + NDIG, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("digits=%3d threads=%d grid=%d: %llu atomics, mismatches vs ascending-lane order = %llu, vs descending = %llu  (%s)\n", NDIG, threads, grid, h[2], h[0],
           h[1], cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    run<1>(384, 296, 4000);
    run<2>(384, 296, 4000);
    run<5>(384, 296, 4000);
    run<16>(384, 296, 4000);
    run<33>(384, 296, 4000);
    run<256>(384, 296, 4000);
    run<256>(128, 1184, 4000);
    run<7>(32, 148, 4000);
    return 0;
}
