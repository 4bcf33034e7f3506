// psac-b200: single-GPU LSD radix sort of (key, value) pairs -- the "(B1,B2) tuple sort" of the
// prefix-doubling loop (reference: include/idxsort.hpp:22-83 -> mxx::sort).
//
// Design (B200-first, not the reference's comparison sample sort):
//   * one up-front histogram kernel counts every digit of every pass in a single read of the keys;
//   * one kernel per 8-bit digit.  A CTA owns a tile, ranks its keys per digit with warp-level
//     match_any multi-split (stable), resolves its global bin offsets with a decoupled look-back over
//     256 per-digit channels (no second pass over the data, no grid-wide sync) and writes the tile
//     out through shared memory so each bin's run leaves as one coalesced burst;
//   * algorithmic HBM traffic per pass = read + write of every key and value once.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace psacb200 {

struct NoVal {};  // keys-only sort

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;

struct RadixPlan {
    int npass;
    int shift[MAX_PASSES];
    int bits[MAX_PASSES];
};

static inline RadixPlan make_radix_plan(int begin_bit, int end_bit) {
    RadixPlan p{};
    int nbits = end_bit - begin_bit;
    if (nbits <= 0) {
        p.npass = 0;
        return p;
    }
    p.npass = (nbits + RADIX_BITS - 1) / RADIX_BITS;
    // spread the bits evenly over the passes (e.g. 42 bits -> 6 passes of 7 bits)
    int base = nbits / p.npass, extra = nbits % p.npass, s = begin_bit;
    for (int i = 0; i < p.npass; ++i) {
        p.shift[i] = s;
        p.bits[i] = base + (i < extra ? 1 : 0);
        s += p.bits[i];
    }
    return p;
}

// ------------------------------------------------------------------ up-front digit histogram
template <typename KeyT>
__device__ __forceinline__ void hist_accumulate(u32 (*sh)[RADIX], const RadixPlan& plan, KeyT k) {
#pragma unroll
    for (int p = 0; p < MAX_PASSES; ++p) {
        if (p < plan.npass) atomicAdd(&sh[p][(u32)(k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u)], 1u);
    }
}

__device__ __forceinline__ void hist_flush(u32 (*sh)[RADIX], const RadixPlan& plan, u64* ghist) {
    for (int e = threadIdx.x; e < plan.npass * RADIX; e += blockDim.x) {
        u32 c = sh[e >> RADIX_BITS][e & (RADIX - 1)];
        if (c) atomicAdd((unsigned long long*)&ghist[e], (unsigned long long)c);
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(512) radix_hist_kernel(const KeyT* __restrict__ keys, size_t n, RadixPlan plan, u64* __restrict__ ghist) {
    __shared__ u32 sh[MAX_PASSES][RADIX];
    for (int e = threadIdx.x; e < MAX_PASSES * RADIX; e += blockDim.x) (&sh[0][0])[e] = 0;
    __syncthreads();
    constexpr int VEC = 16 / sizeof(KeyT);
    const size_t nvec = n / VEC;
    const uint4* kv = reinterpret_cast<const uint4*>(keys);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 q = __ldcs(kv + i);
        if (sizeof(KeyT) == 8) {
            hist_accumulate<u64>(sh, plan, ((u64)q.y << 32) | q.x);
            hist_accumulate<u64>(sh, plan, ((u64)q.w << 32) | q.z);
        } else {
            hist_accumulate<u32>(sh, plan, q.x);
            hist_accumulate<u32>(sh, plan, q.y);
            hist_accumulate<u32>(sh, plan, q.z);
            hist_accumulate<u32>(sh, plan, q.w);
        }
    }
    if (blockIdx.x == 0) {
        for (size_t i = nvec * VEC + threadIdx.x; i < n; i += blockDim.x) hist_accumulate<KeyT>(sh, plan, keys[i]);
    }
    __syncthreads();
    hist_flush(sh, plan, ghist);
}

// exclusive scan of each pass's 256-bin histogram; one CTA of 256 threads per pass
__global__ void __launch_bounds__(RADIX) radix_scan_hist_kernel(const u64* __restrict__ ghist, u64* __restrict__ gbase) {
    __shared__ u64 wtot[RADIX / 32];
    const int p = blockIdx.x, d = threadIdx.x;
    u64 c = ghist[p * RADIX + d];
    u64 inc = warp_inclusive_scan(c, OpSum());
    if ((d & 31) == 31) wtot[d >> 5] = inc;
    __syncthreads();
    u64 pre = 0;
    for (int w = 0; w < (d >> 5); ++w) pre += wtot[w];
    gbase[p * RADIX + d] = pre + inc - c;
}

// ------------------------------------------------------------------ one digit pass
// Shared memory layout of a pass CTA: [tile staging: TILE * max(sizeof key, sizeof value)] [per-warp rank table:
// NW * 256 * {match mask, running position}] [bin_start 256 * u32] [goff 256 * u64] [misc].
template <typename KeyT, typename ValT, int THREADS_, int ITEMS_>
struct PassCfg {
    static constexpr int THREADS = THREADS_;
    static constexpr int ITEMS = ITEMS_;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int NW = THREADS / 32;
    static constexpr bool HAS_VALS = !std::is_same<ValT, NoVal>::value;
    static constexpr size_t ELT = (HAS_VALS && sizeof(ValT) > sizeof(KeyT)) ? sizeof(ValT) : sizeof(KeyT);
    static constexpr size_t OFF_TAB = (size_t)TILE * ELT;
    static constexpr size_t OFF_BIN = OFF_TAB + (size_t)NW * RADIX * 8;
    static constexpr size_t OFF_GOFF = OFF_BIN + RADIX * 4;
    static constexpr size_t OFF_MISC = OFF_GOFF + RADIX * 8;
    static constexpr size_t SMEM = OFF_MISC + 64;
};

// Stable ranking of the warp's keys by digit.  match.any costs ~50 cycles per warp instruction per SM on B200
// (measured, tools/micro_rank.cu) -- more than the whole HBM budget of a key -- so peers are found through a
// shared-memory table instead: every lane ORs its lane bit into mask[d]; one 64-bit read returns {peers, running
// position}; the lowest peer clears the mask and advances the position (~13 cycles per warp instruction).
template <typename KeyT, typename ValT, int THREADS, int ITEMS, bool FULL>
__device__ __forceinline__ void onesweep_tile(unsigned char* smem_raw, const KeyT* __restrict__ kin, KeyT* __restrict__ kout,
                                              const ValT* __restrict__ vin, ValT* __restrict__ vout, const size_t base, const int valid,
                                              const int shift, const u32 mask, const u64* __restrict__ gbase, u64* __restrict__ lookback,
                                              const size_t tile, const u32 epoch) {
    using Cfg = PassCfg<KeyT, ValT, THREADS, ITEMS>;
    constexpr int NW = Cfg::NW;
    KeyT* skeys = reinterpret_cast<KeyT*>(smem_raw);
    uint2* tab = reinterpret_cast<uint2*>(smem_raw + Cfg::OFF_TAB);  // [NW][RADIX] {x: match mask, y: count / position}
    u32* bin_start = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_BIN);
    u64* goff = reinterpret_cast<u64*>(smem_raw + Cfg::OFF_GOFF);
    u32* misc = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_MISC);  // [1..8] warp totals of the digit scan

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int woff = warp * 32 * ITEMS + lane;
    uint2* mytab = tab + warp * RADIX;
    const u32 lane_bit = 1u << lane;
    const u32 lt = lanemask_lt();

    // ---- load keys (and values), warp-striped: item j of lane l is tile element warp*32*ITEMS + j*32 + l
    KeyT key[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int o = woff + j * 32;
        key[j] = (FULL || o < valid) ? ld_stream(kin + base + o) : (KeyT)0;
    }

    // ---- early counts: per-warp digit histogram with fire-and-forget shared atomics
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        if (FULL || (woff + j * 32) < valid) atomicAdd(&mytab[(u32)(key[j] >> shift) & mask].y, 1u);
    }
    __syncthreads();

    // ---- per digit: exclusive offsets of the warps, tile count (published at once for the look-back), bin start
    u32 count = 0;
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const u32 c = tab[w * RADIX + tid].y;
            tab[w * RADIX + tid].y = count;
            count += c;
        }
        if (tile > 0) st_relaxed(lookback + tile * RADIX + tid, lb_pack((u64)count, epoch, LB_AGGREGATE));
    }
    const u32 inc = warp_inclusive_sum_u32(count);
    if (tid < RADIX && lane == 31) misc[1 + warp] = inc;
    __syncthreads();
    if (tid < RADIX) {
        u32 pre = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w) pre += (w < warp) ? misc[1 + w] : 0u;
        const u32 bstart = pre + inc - count;
        bin_start[tid] = bstart;
#pragma unroll
        for (int w = 0; w < NW; ++w) tab[w * RADIX + tid].y += bstart;  // position of the warp's first key of this digit
    }
    __syncthreads();

    // ---- stable ranking -> final position of every key inside the tile
    u16 pos[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const bool ok = FULL || (woff + j * 32) < valid;
        const u32 d = (u32)(key[j] >> shift) & mask;
        if (ok) atomicOr(&mytab[d].x, lane_bit);
        __syncwarp();
        const uint2 e = mytab[d];
        __syncwarp();
        if (ok && (e.x & lt) == 0) mytab[d] = make_uint2(0u, e.y + __popc(e.x));
        pos[j] = (u16)(e.y + __popc(e.x & lt));
        __syncwarp();
    }

    // values: issue the global loads now, they overlap the key scatter, the look-back and the key write-out
    ValT val[Cfg::HAS_VALS ? ITEMS : 1];
    if constexpr (Cfg::HAS_VALS) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int o = woff + j * 32;
            if (FULL || o < valid) val[j] = ld_stream(vin + base + o);
        }
    }

    // ---- scatter keys into shared memory in bin order (tile staging is free: keys live in registers)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        if (FULL || (woff + j * 32) < valid) skeys[pos[j]] = key[j];
    }

    // ---- decoupled look-back, one channel per digit; the predecessors published their counts long ago
    if (tid < RADIX) {
        u64 excl = 0;
        if (tile == 0) {
            st_relaxed(lookback + tid, lb_pack((u64)count, epoch, LB_INCLUSIVE));
        } else {
            size_t t = tile;
            while (true) {
                --t;
                u64 w, st;
                do {
                    w = ld_relaxed(lookback + t * RADIX + tid);
                    st = lb_state(w, epoch);
                } while (st == LB_NONE);
                excl += lb_payload(w);
                if (st == LB_INCLUSIVE) break;
            }
            st_relaxed(lookback + tile * RADIX + tid, lb_pack(excl + count, epoch, LB_INCLUSIVE));
        }
        goff[tid] = gbase[tid] + excl - (u64)bin_start[tid];
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive shared positions of one bin are consecutive in global memory
    u8 dig[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int s = i * THREADS + tid;
        if (FULL || s < valid) {
            const KeyT k = skeys[s];
            const u32 d = (u32)(k >> shift) & mask;
            dig[i] = (u8)d;
            kout[goff[d] + (u64)s] = k;
        }
    }
    if constexpr (Cfg::HAS_VALS) {
        ValT* svals = reinterpret_cast<ValT*>(smem_raw);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (FULL || (woff + j * 32) < valid) svals[pos[j]] = val[j];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int s = i * THREADS + tid;
            if (FULL || s < valid) vout[goff[dig[i]] + (u64)s] = svals[s];
        }
    }
}

template <typename KeyT, typename ValT, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, 2) onesweep_pass_kernel(const KeyT* __restrict__ kin, KeyT* __restrict__ kout,
                                                               const ValT* __restrict__ vin, ValT* __restrict__ vout, size_t n, int shift,
                                                               int bits, const u64* __restrict__ gbase, u32* __restrict__ tile_counter,
                                                               u64* __restrict__ lookback, u32 epoch) {
    using Cfg = PassCfg<KeyT, ValT, THREADS, ITEMS>;
    constexpr int TILE = Cfg::TILE;
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "need one thread per digit");
    static_assert(TILE < 65536, "tile positions are kept in 16 bits");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* misc = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_MISC);
    uint2* tab = reinterpret_cast<uint2*>(smem_raw + Cfg::OFF_TAB);
    if (threadIdx.x == 0) misc[0] = atomicAdd(tile_counter, 1u);
    for (int e = threadIdx.x; e < Cfg::NW * RADIX; e += THREADS) tab[e] = make_uint2(0u, 0u);
    __syncthreads();
    const size_t tile = misc[0];
    const size_t base = tile * (size_t)TILE;
    const u32 mask = (1u << bits) - 1u;
    if (n - base >= (size_t)TILE)
        onesweep_tile<KeyT, ValT, THREADS, ITEMS, true>(smem_raw, kin, kout, vin, vout, base, TILE, shift, mask, gbase, lookback, tile, epoch);
    else
        onesweep_tile<KeyT, ValT, THREADS, ITEMS, false>(smem_raw, kin, kout, vin, vout, base, (int)(n - base), shift, mask, gbase, lookback, tile,
                                                       epoch);
}

// ------------------------------------------------------------------ host driver
template <typename KeyT, typename ValT>
struct SortTuning {
    static constexpr int THREADS = 384;
    static constexpr int ITEMS = (sizeof(KeyT) + (std::is_same<ValT, NoVal>::value ? 0 : sizeof(ValT)) >= 16) ? 12 : 16;
};

struct RadixWorkspace {
    u64* ghist = nullptr;      // [MAX_PASSES][RADIX]
    u64* gbase = nullptr;      // [MAX_PASSES][RADIX]
    u32* counters = nullptr;   // [MAX_PASSES]
    u64* lookback = nullptr;   // [max_tiles][RADIX]
    size_t lookback_bytes = 0;
    static size_t small_bytes() { return 2 * MAX_PASSES * RADIX * sizeof(u64) + 64 * sizeof(u32); }
    template <typename KeyT, typename ValT>
    static size_t lookback_bytes_for(size_t n) {
        using T = SortTuning<KeyT, ValT>;
        return div_up(n ? n : 1, (size_t)T::THREADS * T::ITEMS) * RADIX * sizeof(u64);
    }
};

// Sorts n pairs by key bits [begin_bit, end_bit).  Ping-pongs between (keys, vals) and (keys_alt, vals_alt);
// returns true when the sorted data ended up in the *_alt buffers.
template <typename KeyT, typename ValT>
bool radix_sort_pairs(const RadixWorkspace& ws, KeyT* keys, KeyT* keys_alt, ValT* vals, ValT* vals_alt, size_t n, int begin_bit, int end_bit,
                      cudaStream_t stream, int sm_count, RadixPlan* plan_out = nullptr, uint64_t* launches = nullptr,
                      cudaEvent_t ev_hist_done = nullptr, cudaEvent_t ev_passes_begin = nullptr) {
    using T = SortTuning<KeyT, ValT>;
    using Cfg = PassCfg<KeyT, ValT, T::THREADS, T::ITEMS>;
    RadixPlan plan = make_radix_plan(begin_bit, end_bit);
    if (plan_out) *plan_out = plan;
    if (n == 0 || plan.npass == 0) return false;
    const size_t tiles = div_up(n, (size_t)Cfg::TILE);
    if (tiles * RADIX * sizeof(u64) > ws.lookback_bytes) throw std::string("radix_sort_pairs: look-back workspace too small");
    auto kern = onesweep_pass_kernel<KeyT, ValT, T::THREADS, T::ITEMS>;
    static bool attr_set = false;
    if (!attr_set) {
        PSAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    {
        PSAC_CUDA(cudaMemsetAsync(ws.ghist, 0, MAX_PASSES * RADIX * sizeof(u64), stream));
        size_t want = div_up(n, (size_t)512 * 32);
        int grid = (int)(want < (size_t)sm_count * 4 ? (want ? want : 1) : (size_t)sm_count * 4);
        radix_hist_kernel<KeyT><<<grid, 512, 0, stream>>>(keys, n, plan, ws.ghist);
    }
    radix_scan_hist_kernel<<<plan.npass, RADIX, 0, stream>>>(ws.ghist, ws.gbase);
    if (launches) *launches = 2 + (uint64_t)plan.npass;
    if (ev_hist_done) cudaEventRecord(ev_hist_done, stream);
    if (ev_passes_begin) cudaEventRecord(ev_passes_begin, stream);
    PSAC_CUDA(cudaMemsetAsync(ws.counters, 0, MAX_PASSES * sizeof(u32), stream));
    PSAC_CUDA(cudaMemsetAsync(ws.lookback, 0, tiles * RADIX * sizeof(u64), stream));
    bool in_alt = false;
    for (int p = 0; p < plan.npass; ++p) {
        KeyT* ki = in_alt ? keys_alt : keys;
        KeyT* ko = in_alt ? keys : keys_alt;
        ValT* vi = in_alt ? vals_alt : vals;
        ValT* vo = in_alt ? vals : vals_alt;
        kern<<<(unsigned)tiles, T::THREADS, Cfg::SMEM, stream>>>(ki, ko, vi, vo, n, plan.shift[p], plan.bits[p], ws.gbase + p * RADIX,
                                                                 ws.counters + p, ws.lookback, (u32)(p + 1));
        in_alt = !in_alt;
    }
    PSAC_CUDA(cudaGetLastError());
    return in_alt;
}

}  // namespace psacb200
