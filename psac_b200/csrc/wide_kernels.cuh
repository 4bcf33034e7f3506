// psac-b200: texts over wide characters (2 or 4 bytes per character).
//
// Reference: suffix_array<int, index_t, LCP> orders the characters by value through int_alphabet
// (include/alphabet.hpp:355-513; test/test_psac.cpp:277-304 "IntAlphabetMiss": SA of {128, 3, 12345678, ...} = SA of
// "mississippi").  SA / ISA / LCP depend only on the order and equality of the characters, so the text is reduced on the
// device to one byte per character = 1 + rank of its value among the values that occur (at most 255 distinct values),
// and the byte construction runs on that.  Distinct values: a hash set per CTA in shared memory, merged into one global
// set; ranks: binary search in the sorted table.
#pragma once
#include "common.cuh"

namespace psacb200 {

constexpr int WIDE_SLOTS = 1024;      // slots of the CTA-local and of the global hash set (open addressing, <= 255 live entries)
constexpr int WIDE_MAX_DISTINCT = 255;

// value -> unsigned key in value order (signed types: flip the sign bit)
template <typename T>
__device__ __forceinline__ u32 wide_key(T v, u32 flip) {
    return ((u32)(typename std::conditional<sizeof(T) == 2, unsigned short, u32>::type)v) ^ flip;
}

// inserts (1 << 32 | key) into a set of WIDE_SLOTS 64-bit slots (0 = empty); returns false when the set is full
__device__ __forceinline__ bool wide_insert(unsigned long long* tab, u32 key, bool* fresh) {
    const unsigned long long e = (1ull << 32) | key;
    u32 h = (key * 2654435761u) >> 22;
    for (int probe = 0; probe < WIDE_SLOTS; ++probe) {
        const unsigned long long cur = tab[h];
        if (cur == e) return true;
        if (cur == 0) {
            const unsigned long long old = atomicCAS(&tab[h], 0ull, e);
            if (old == 0) {
                *fresh = true;
                return true;
            }
            if (old == e) return true;
        }
        h = (h + 1) & (WIDE_SLOTS - 1);
    }
    return false;
}

// meta[0] = number of distinct values in gtab, meta[1] = overflow flag
template <typename T>
__global__ void __launch_bounds__(256) wide_distinct_kernel(const T* __restrict__ text, u64 n, u32 flip, unsigned long long* __restrict__ gtab,
                                                            unsigned int* __restrict__ meta) {
    __shared__ unsigned long long s_tab[WIDE_SLOTS];
    __shared__ unsigned int s_cnt;
    for (int e = threadIdx.x; e < WIDE_SLOTS; e += blockDim.x) s_tab[e] = 0;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    u32 last = 0;
    bool have_last = false;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 k = wide_key<T>(text[i], flip);
        if (have_last && k == last) continue;
        last = k;
        have_last = true;
        if (s_cnt > WIDE_MAX_DISTINCT) break;  // (too many: the global flag is raised below)
        bool fresh = false;
        if (!wide_insert(s_tab, k, &fresh)) break;
        if (fresh) atomicAdd(&s_cnt, 1u);
    }
    __syncthreads();
    if (s_cnt > WIDE_MAX_DISTINCT) {
        if (threadIdx.x == 0) meta[1] = 1u;
        return;
    }
    for (int e = threadIdx.x; e < WIDE_SLOTS; e += blockDim.x) {
        const unsigned long long v = s_tab[e];
        if (v == 0) continue;
        if (meta[0] > WIDE_MAX_DISTINCT) {
            meta[1] = 1u;
            break;
        }
        bool fresh = false;
        if (!wide_insert(gtab, (u32)v, &fresh)) {
            meta[1] = 1u;
            break;
        }
        if (fresh) atomicAdd(&meta[0], 1u);
    }
}

// out[i] = 1 + rank of text[i] in the sorted table of the `cnt` keys that occur
template <typename T>
__global__ void __launch_bounds__(256) wide_map_kernel(const T* __restrict__ text, u64 n, u32 flip, const u32* __restrict__ table, int cnt,
                                                       u8* __restrict__ out) {
    __shared__ u32 s_t[256];
    s_t[threadIdx.x] = (int)threadIdx.x < cnt ? table[threadIdx.x] : 0xffffffffu;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 k = wide_key<T>(text[i], flip);
        int lo = 0, hi = cnt;  // first entry >= k (k occurs in the table)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_t[mid] < k)
                lo = mid + 1;
            else
                hi = mid;
        }
        out[i] = (u8)(lo + 1);
    }
}

}  // namespace psacb200
