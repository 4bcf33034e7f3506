// Micro-benchmark: throughput per SM of the warp-level primitives a radix-rank loop can be built from.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro_rank tools/micro_rank.cu && ./tools/micro_rank
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 2048
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int MODE>
__global__ void k(unsigned* out, unsigned seed, long long* cycles) {
    __shared__ unsigned cnt[32][256];
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) (&cnt[0][0])[i] = 0;
    __syncthreads();
    unsigned x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
        x = x * 1664525u + 1013904223u;
        unsigned d = (x >> 24) & 0xff;
        if (MODE == 0) {  // match.any only
            acc += __match_any_sync(0xffffffffu, d);
        } else if (MODE == 1) {  // 8 ballots
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1);
                peers &= ((d >> b) & 1) ? m : ~m;
            }
            acc += peers;
        } else if (MODE == 2) {  // full rank step, match + leader RMW + shfl (current kernel)
            unsigned peers = __match_any_sync(0xffffffffu, d);
            int leader = __ffs(peers) - 1;
            unsigned before = 0;
            if (lane == leader) { before = cnt[warp][d]; cnt[warp][d] = before + __popc(peers); }
            before = __shfl_sync(0xffffffffu, before, leader);
            acc += before + __popc(peers & lanemask_lt());
            __syncwarp();
        } else if (MODE == 3) {  // full rank step, ballots + all-read + leader write
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1);
                peers &= ((d >> b) & 1) ? m : ~m;
            }
            unsigned before = cnt[warp][d];
            __syncwarp();
            if ((peers & lanemask_lt()) == 0) cnt[warp][d] = before + __popc(peers);
            acc += before + __popc(peers & lanemask_lt());
            __syncwarp();
        } else if (MODE == 4) {  // smem atomicAdd with return (unstable rank)
            acc += atomicAdd(&cnt[warp][d], 1u);
        } else if (MODE == 5) {  // match + all-read + leader write (no shfl)
            unsigned peers = __match_any_sync(0xffffffffu, d);
            unsigned before = cnt[warp][d];
            __syncwarp();
            if ((peers & lanemask_lt()) == 0) cnt[warp][d] = before + __popc(peers);
            acc += before + __popc(peers & lanemask_lt());
            __syncwarp();
        } else if (MODE == 6) {  // match + leader atomicAdd + shfl (CUB style)
            unsigned peers = __match_any_sync(0xffffffffu, d);
            int leader = 31 - __clz(peers);
            unsigned pc = __popc(peers & (lanemask_lt() | (1u << lane)));
            unsigned before = 0;
            if (lane == leader) before = atomicAdd(&cnt[warp][d], pc);
            before = __shfl_sync(0xffffffffu, before, leader);
            acc += before + pc - 1;
        } else if (MODE == 7) {  // smem atomicAdd no return (histogram)
            atomicAdd(&cnt[warp][d], 1u);
        } else if (MODE == 8) {  // rank: atomicOr match table {mask,count} + LDS.64 + leader STS.64
            uint2* tab = reinterpret_cast<uint2*>(&cnt[0][0]) + (warp & 15) * 256;  // 2 KB per warp (only 16 warps fit in the 32 KB array)
            atomicOr(&tab[d].x, 1u << lane);
            __syncwarp();
            uint2 e = tab[d];
            __syncwarp();
            if ((e.x & lanemask_lt()) == 0) tab[d] = make_uint2(0u, e.y + __popc(e.x));
            acc += e.y + __popc(e.x & lanemask_lt());
            __syncwarp();
        } else if (MODE == 9) {  // two interleaved chains of MODE 8 (separate tables), 2 keys per iteration
            uint2* tabA = reinterpret_cast<uint2*>(&cnt[0][0]) + (warp % 8) * 512;
            uint2* tabB = tabA + 256;
            unsigned d2 = (x >> 16) & 0xff;
            atomicOr(&tabA[d].x, 1u << lane);
            atomicOr(&tabB[d2].x, 1u << lane);
            __syncwarp();
            uint2 e = tabA[d];
            uint2 f = tabB[d2];
            __syncwarp();
            if ((e.x & lanemask_lt()) == 0) tabA[d] = make_uint2(0u, e.y + __popc(e.x));
            if ((f.x & lanemask_lt()) == 0) tabB[d2] = make_uint2(0u, f.y + __popc(f.x));
            acc += e.y + __popc(e.x & lanemask_lt()) + f.y + __popc(f.x & lanemask_lt());
            __syncwarp();
        } else if (MODE == 10) {  // split tables: atomicOr mask[d] + LDS mask + LDS count + leader STS x2
            unsigned* mask = &cnt[0][0] + (warp & 15) * 512;
            unsigned* cn = mask + 256;
            atomicOr(&mask[d], 1u << lane);
            __syncwarp();
            unsigned m = mask[d];
            unsigned c = cn[d];
            __syncwarp();
            if ((m & lanemask_lt()) == 0) { mask[d] = 0u; cn[d] = c + __popc(m); }
            acc += c + __popc(m & lanemask_lt());
            __syncwarp();
        } else if (MODE == 11) {  // ballots on 8 bits, peers only, then count via per-warp table packed 2x16 (one LDS + leader STS)
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1);
                peers &= ((d >> b) & 1) ? m : ~m;
            }
            unsigned short* tab16 = reinterpret_cast<unsigned short*>(&cnt[0][0]) + warp * 256;
            unsigned before = tab16[d];
            __syncwarp();
            if ((peers & lanemask_lt()) == 0) tab16[d] = (unsigned short)(before + __popc(peers));
            acc += before + __popc(peers & lanemask_lt());
            __syncwarp();
        } else if (MODE == 12) {  // unstable: CTA-wide table, atomicAdd with return
            acc += atomicAdd(&cnt[0][d], 1u);
        } else if (MODE == 13) {  // plain LDS gather + STS scatter with random digit index (cost of one table lookup)
            unsigned c = cnt[warp][d];
            acc += c;
            cnt[warp][(d + lane) & 255] = c + 1;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + cnt[warp][lane];
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm) {
    int sms = 148;
    int grid = sms * ctas_per_sm;
    unsigned* out; long long* cyc;
    cudaMalloc(&out, grid * threads * 4); cudaMalloc(&cyc, grid * 8);
    k<MODE><<<grid, threads>>>(out, 1, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, threads>>>(out, 2, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[4096]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    double warp_ops_per_sm = (double)ITERS * (threads / 32) * ctas_per_sm;
    printf("%-40s thr=%4d cta/sm=%d  cycles/CTA=%9.0f  => %.2f cycles per warp-op per SM  (%.3f ms)\n", name, threads, ctas_per_sm, avg, avg / warp_ops_per_sm * 1.0, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int c = 1; c <= 2; ++c) {
        run<0>("match.any", 384, c);
        run<1>("8 ballots", 384, c);
        run<2>("rank: match+leader RMW+shfl (v1)", 384, c);
        run<3>("rank: ballots+allread+leaderwrite", 384, c);
        run<5>("rank: match+allread+leaderwrite", 384, c);
        run<6>("rank: match+leader atomicAdd+shfl (CUB)", 384, c);
        run<4>("smem atomicAdd w/ return", 384, c);
        run<7>("smem atomicAdd no return", 384, c);
        run<8>("rank: atomicOr table + LDS64 + STS64", 384, c);
        run<9>("rank: 2 chains of the above (2 keys/iter)", 256, c);
        run<10>("rank: atomicOr split tables", 384, c);
        run<11>("rank: ballots + u16 table", 384, c);
        run<12>("unstable: CTA atomicAdd w/ return", 384, c);
        run<13>("LDS gather + STS scatter", 384, c);
    }
    return 0;
}
