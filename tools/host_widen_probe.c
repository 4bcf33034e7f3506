// Host-side probe: how fast can T threads widen u32 -> u64 into a second buffer (the host half of a "32-bit over PCIe,
// widen on the host" output path)?  gcc -O3 -march=x86-64-v3 -pthread tools/host_widen_probe.c -o tools/host_widen_probe
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
typedef struct { const uint32_t* s; uint64_t* d; size_t n; int nt; } job;
static void widen(const uint32_t* s, uint64_t* d, size_t n, int nt) {
    size_t i = 0;
    if (nt) {
        for (; i + 8 <= n; i += 8) {
            __m256i v = _mm256_loadu_si256((const __m256i*)(s + i));
            __m256i lo = _mm256_cvtepu32_epi64(_mm256_castsi256_si128(v));
            __m256i hi = _mm256_cvtepu32_epi64(_mm256_extracti128_si256(v, 1));
            _mm256_stream_si256((__m256i*)(d + i), lo);
            _mm256_stream_si256((__m256i*)(d + i + 4), hi);
        }
    }
    for (; i < n; ++i) d[i] = s[i];
}
static void* run(void* p) { job* j = (job*)p; widen(j->s, j->d, j->n, j->nt); return 0; }
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(void) {
    const size_t n = (size_t)1 << 28;
    uint32_t* s = aligned_alloc(64, n * 4);
    uint64_t* d = aligned_alloc(64, n * 8);
    for (size_t i = 0; i < n; ++i) s[i] = (uint32_t)i;
    memset(d, 1, n * 8);
    for (int nt = 0; nt < 2; ++nt)
        for (int T = 1; T <= 64; T *= 2) {
            pthread_t th[64]; job jb[64];
            double t0 = now();
            for (int t = 0; t < T; ++t) {
                size_t a = n / T * t, b = (t == T - 1) ? n : n / T * (t + 1);
                a &= ~(size_t)7;
                if (t != T - 1) b &= ~(size_t)7;
                jb[t] = (job){s + a, d + a, b - a, nt};
                pthread_create(&th[t], 0, run, &jb[t]);
            }
            for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
            double dt = now() - t0;
            printf("%s threads=%2d  %.1f ms  write %.1f GB/s\n", nt ? "stream" : "plain ", T, dt * 1e3, n * 8 / dt / 1e9);
        }
    return (int)(d[12345] & 1);
}
