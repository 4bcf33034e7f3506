// psac-b200: single-GPU radix sort of (key, value) pairs -- the "(B1,B2) tuple sort" of the
// prefix-doubling loop (reference: include/idxsort.hpp:22-83 -> mxx::sort).
//
// Design (B200-first, not the reference's comparison sample sort):
//   * one 8-bit digit per pass.  A pass is: per-tile digit histogram (reads the keys only) -> scan of the per-tile
//     counts (tiny kernels) -> scatter kernel.  A scatter CTA owns a tile, ranks its keys with ONE shared-memory
//     atomic per key (see "ranking" below), adds the tile's global bin offsets and writes the tile out through
//     shared memory so each bin's run leaves as one coalesced burst.
//     (Measured first: the usual single-kernel "onesweep" with a decoupled look-back.  On B200 ~450 tiles are resident
//     and a dependent L2 load under full DRAM load costs ~1 us, so 54 % of a tile's cycles were look-back waits
//     (profiles/r1_bench_pass.txt); re-reading 4 bytes per key for the histogram is cheaper than that.)
//   * generic pairs (radix_sort_pairs) and 64-bit carried keys (radix_sort_suffixes) are sorted LSD;
//   * the first sort of a construction with <= 40 key bits (BASELINE configs[1]: 20 DNA characters) goes TOP DIGIT FIRST
//     (radix_sort_suffixes_msd): pass 1 reads the packed text instead of a key array (k-mer generation fused) and
//     partitions by the top digit into 256 tile-aligned segments, the other digits are sorted LSD inside the segments,
//     so only the low 32 key bits travel with the 32-bit suffix index -- the top digit is implied by the segment;
//   * algorithmic HBM traffic per pass = read + write of every carried key and value once (the histogram's re-read of
//     the keys is overhead and counted as such in DESIGN.md).
//
// Ranking.  A stable rank needs, for every key, the number of earlier keys of the tile with the same digit.
// On sm_100a the lanes of one ATOMS.ADD warp instruction that hit the same shared-memory word are applied in
// ascending lane order, and successive ATOMS of a warp are applied in program order (profiles/r1_atoms_order.txt:
// 0 mismatches in 3.2e9 atomics; re-verified by atoms_order_selftest at every engine creation).  So with one
// 256-entry counter table per warp and warp-striped items, the value returned by atomicAdd(&tab[warp][digit], 1)
// IS the key's stable rank inside its warp.  That is 2.6 cycles per warp instruction per SM against 13.4 for
// the atomicOr match-table sequence it replaces and 47-62 for match.any (profiles/r1_micro_rank.txt).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace psacb200 {

// optional per-phase cycle accounting of the pass kernel (tools/bench_pass.cu defines PSAC_PHASE_PROFILE)
#ifdef PSAC_PHASE_PROFILE
__device__ unsigned long long g_phase_cycles[16];
#define PSAC_PHASE(i)                                                          \
    do {                                                                       \
        if (threadIdx.x == 0) {                                                \
            const long long _c = clock64();                                    \
            atomicAdd(&g_phase_cycles[i], (unsigned long long)(_c - _phase_t)); \
            _phase_t = _c;                                                     \
        }                                                                      \
    } while (0)
#define PSAC_PHASE_BEGIN() long long _phase_t = clock64()
#else
#define PSAC_PHASE(i) do { } while (0)
#define PSAC_PHASE_BEGIN() do { } while (0)
#endif

struct NoVal {};  // keys-only sort

// sources may offer hist_digit(g): the digit of element g without building the complete staged key (histogram pre-pass)
template <class Src, class = void>
struct has_hist_digit : std::false_type {};
template <class Src>
struct has_hist_digit<Src, std::void_t<decltype(std::declval<const Src&>().hist_digit((size_t)0))>> : std::true_type {};

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;
constexpr int SCAN_CHUNK = 256;  // tiles per chunk of the two-level scan of the per-tile digit counts

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

// ------------------------------------------------------------------ key sources of a pass
// A source provides the staged key of global element g, its digit, the key to write out, the value and (optionally)
// an auxiliary byte that travels with the pair.
template <typename KeyT, typename ValT>
struct ArraySrc {
    using Stage = KeyT;
    using Out = KeyT;
    static constexpr bool FROM_TEXT = false;
    static constexpr bool PEER = false;  // outputs go to one array pair (see OwnerSrc in sharded.cuh for per-digit outputs)
    const KeyT* __restrict__ kin;
    const ValT* __restrict__ vin;
    const u8* __restrict__ ain;  // auxiliary bytes (null when the pass carries none)
    int shift;
    u32 mask;
    KeyT sub;                    // subtracted from the key before the digit is taken (window partitions of a shard's block)
    __device__ __forceinline__ Stage load_key(size_t g) const { return ld_stream(kin + g); }
    __device__ __forceinline__ u32 digit(Stage k) const { return (u32)((k - sub) >> shift) & mask; }
    __device__ __forceinline__ Out out_key(Stage k) const { return k; }
    __device__ __forceinline__ ValT load_val(size_t g) const { return ld_stream(vin + g); }
    __device__ __forceinline__ u8 load_aux(size_t g, Stage) const { return ld_stream(ain + g); }
};

// (suffix, position) pairs of the SA -> ISA step: the key is the suffix index read from the sorted array, the value its
// position g in that array -- generated, not read
template <typename KeyT, typename ValT>
struct PosSrc {
    using Stage = KeyT;
    using Out = KeyT;
    static constexpr bool FROM_TEXT = false;
    static constexpr bool PEER = false;
    const KeyT* __restrict__ kin;
    int shift;
    u32 mask;
    __device__ __forceinline__ Stage load_key(size_t g) const { return ld_stream(kin + g); }
    __device__ __forceinline__ u32 digit(Stage k) const { return (u32)(k >> shift) & mask; }
    __device__ __forceinline__ Out out_key(Stage k) const { return k; }
    __device__ __forceinline__ ValT load_val(size_t g) const { return (ValT)g; }
    __device__ __forceinline__ u8 load_aux(size_t, Stage) const { return 0; }
};

// First pass of a construction (reference a4 k-mer generation, include/kmer.hpp:119-224, fused into the sort):
// element g is suffix idx(g); its key is the first kbits stream bits of that suffix; the digit is the low `drop` bits
// and the key carried on is the rest; the dropped digit travels on as the auxiliary byte (the resolve step needs the
// complete key).  With drop == 0 the whole key is carried (64-bit carried keys) and there is no auxiliary byte.
// Initial order: the T = min(n, C-1) suffixes that run past the end of the text come FIRST, shortest first, then
// suffixes 0..n-T-1.  The sort is stable, so among equal keys the suffixes that hit the end sort in front and by
// increasing length -- exactly the order the reference obtains from its 0 sentinel code (include/alphabet.hpp:157-164:
// code 0 is reserved for "past the end").
template <typename OutKeyT, typename IdxT>
struct TextSrc {
    using Stage = u64;
    using Out = OutKeyT;
    static constexpr bool FROM_TEXT = true;
    static constexpr bool PEER = false;
    const u64* __restrict__ stream;
    u64 n, T;
    int lbits, kbits, drop;  // drop: low key bits removed from the key that is written out
    u32 mask;
    int dshift;              // the digit of this pass is (key >> dshift) & mask
    __device__ __forceinline__ u64 idx(size_t g) const { return (g < T) ? (n - 1 - g) : (g - T); }
    __device__ __forceinline__ Stage load_key(size_t g) const { return stream_extract(stream, idx(g), lbits, kbits); }
    // key of element g from a shared-memory copy of the stream words [w0, w0 + nw) (elements g >= T only)
    __device__ __forceinline__ Stage load_key_window(const u64* __restrict__ win, u64 w0, size_t g) const {
        const u64 bit = (g - T) * (u64)lbits;
        const u64 w = (bit >> 6) - w0;
        const unsigned o = (unsigned)(bit & 63);
        const u64 hi = win[w], lo = win[w + 1];
        const u64 v = o ? ((hi << o) | (lo >> (64 - o))) : hi;
        return v >> (64 - kbits);
    }
    __device__ __forceinline__ u32 digit(Stage k) const { return (u32)(k >> dshift) & mask; }
    __device__ __forceinline__ Out out_key(Stage k) const { return (Out)(k >> drop); }
    __device__ __forceinline__ IdxT load_val(size_t g) const { return (IdxT)idx(g); }
    __device__ __forceinline__ u8 load_aux(size_t, Stage k) const { return (u8)((u32)(k >> dshift) & mask); }
};

// ------------------------------------------------------------------ one digit pass
// Shared memory layout of a pass CTA: [tile staging] [aux staging TILE bytes, if any] [per-warp counter tables
// NW * 256 * u32] [bin_start 256 * u32] [goff 256 * u64] [misc].  32-bit keys with 32-bit values are staged PACKED as
// one 64-bit word per pair (one scatter, one read-back); other widths stage the keys, then the values.
template <class Src, typename ValT, int THREADS_, int ITEMS_, bool HAS_AUX_>
struct PassCfg {
    using Stage = typename Src::Stage;
    using Out = typename Src::Out;
    static constexpr int THREADS = THREADS_;
    static constexpr int ITEMS = ITEMS_;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int NW = THREADS / 32;
    static constexpr bool HAS_VALS = !std::is_same<ValT, NoVal>::value;
    static constexpr bool HAS_AUX = HAS_AUX_;
    static constexpr bool PACKED = HAS_VALS && sizeof(Out) == 4 && sizeof(ValT) == 4;
    // a packed text pass no longer has the digit in the staged pair: it is read back from the aux byte
    static_assert(!(PACKED && Src::FROM_TEXT) || HAS_AUX, "packed first pass needs the auxiliary byte");
    static constexpr size_t KEYSZ = PACKED ? 8 : sizeof(Stage);
    static constexpr size_t ELT = (!PACKED && HAS_VALS && sizeof(ValT) > KEYSZ) ? sizeof(ValT) : KEYSZ;
    static constexpr size_t OFF_AUX = (size_t)TILE * ELT;
    static constexpr size_t OFF_TAB = OFF_AUX + (HAS_AUX ? (size_t)TILE : 0);
    static constexpr size_t OFF_BIN = OFF_TAB + (size_t)NW * RADIX * 4;
    static constexpr size_t OFF_GOFF = OFF_BIN + RADIX * 4;
    static constexpr size_t OFF_MISC = OFF_GOFF + RADIX * 8;
    static constexpr size_t SMEM = OFF_MISC + 64;
    static_assert(OFF_TAB % 8 == 0, "table alignment");
};

// SAFE = ranking from documented warp primitives only (match.any peers + one atomic per group of equal digits): the
// path taken when the ATOMS lane-order self-test fails at engine creation, or with PSACB200_SAFE_RANK=1.
template <class Cfg, class Src, typename ValT, bool FULL, bool SAFE>
__device__ __forceinline__ void scatter_tile(unsigned char* smem_raw, const Src& src, typename Src::Out* __restrict__ kout, ValT* __restrict__ vout,
                                              u8* __restrict__ aout, const size_t base, const int valid, const u64* __restrict__ gbase,
                                              const u64* __restrict__ chunk_base, const u32* __restrict__ tile_excl) {
    using Stage = typename Src::Stage;
    using Out = typename Src::Out;
    constexpr int NW = Cfg::NW, ITEMS = Cfg::ITEMS, THREADS = Cfg::THREADS;
    u8* saux = smem_raw + Cfg::OFF_AUX;
    u32* tab = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_TAB);  // [NW][RADIX] counts, then running positions
    u32* bin_start = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_BIN);
    u64* goff = reinterpret_cast<u64*>(smem_raw + Cfg::OFF_GOFF);
    u32* misc = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_MISC);  // [1..8] warp totals of the digit scan

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int woff = warp * 32 * ITEMS + lane;
    u32* mytab = tab + warp * RADIX;

    PSAC_PHASE_BEGIN();
    // ---- load keys, warp-striped: item j of lane l is tile element warp*32*ITEMS + j*32 + l
    Stage key[ITEMS];
    bool windowed = false;
    if constexpr (Src::FROM_TEXT) {
        // the tile's suffixes are consecutive: copy the stream words they touch into shared memory once (the staging
        // area is free until the scatter) and cut the keys out of that window instead of two global loads per key
        if (base >= src.T) {
            windowed = true;
            u64* win = reinterpret_cast<u64*>(smem_raw);
            const u64 w0 = ((base - src.T) * (u64)src.lbits) >> 6;
            const int nw = (int)((((u64)valid * src.lbits + src.kbits + 63) >> 6) + 2);
            for (int e = tid; e < nw; e += THREADS) win[e] = __ldg(src.stream + w0 + e);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const int o = woff + j * 32;
                key[j] = (FULL || o < valid) ? src.load_key_window(win, w0, base + o) : (Stage)0;
            }
            __syncthreads();  // the window is overwritten by the scatter later; all keys are in registers now
        }
    }
    if (!windowed) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int o = woff + j * 32;
            key[j] = (FULL || o < valid) ? src.load_key(base + o) : (Stage)0;
        }
    }
    // ---- count the warp's digits (shared-memory reductions, no return value)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        if (FULL || (woff + j * 32) < valid) atomicAdd(&mytab[src.digit(key[j])], 1u);
    }
    // values / aux bytes: issue the global loads now, they land while the counts are scanned
    ValT val[Cfg::HAS_VALS ? ITEMS : 1];
    u8 aux[Cfg::HAS_AUX ? ITEMS : 1];
    if constexpr (Cfg::HAS_VALS) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int o = woff + j * 32;
            if (FULL || o < valid) val[j] = src.load_val(base + o);
        }
    }
    if constexpr (Cfg::HAS_AUX) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int o = woff + j * 32;
            if (FULL || o < valid) aux[j] = src.load_aux(base + o, key[j]);
        }
    }
    PSAC_PHASE(0);  // key loads + counting
    __syncthreads();
    PSAC_PHASE(1);  // barrier 1

    // ---- per digit: exclusive offsets of the warps, tile count, bin start
    u32 count = 0;
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const u32 c = tab[w * RADIX + tid];
            tab[w * RADIX + tid] = count;
            count += c;
        }
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
        for (int w = 0; w < NW; ++w) tab[w * RADIX + tid] += bstart;  // position of the warp's first key of this digit
    }
    __syncthreads();
    PSAC_PHASE(2);  // digit scan (2 barriers)

    // ---- stable rank = one ATOMS.ADD with return per key on the warp's running positions (lanes of one instruction
    //      apply in ascending lane order, instructions of a warp in program order -- see the header comment), then
    //      scatter into shared memory in bin order
    u32 pos[Cfg::PACKED ? 1 : ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const bool act = FULL || (woff + j * 32) < valid;
        u32 p = 0;
        if constexpr (SAFE) {
            const u32 am = __ballot_sync(0xffffffffu, act);
            if (act) {
                const u32 d = src.digit(key[j]);
                const u32 peers = __match_any_sync(am, d);
                const int leader = 31 - __clz(peers);
                u32 b0 = 0;
                if (lane == leader) b0 = atomicAdd(&mytab[d], (u32)__popc(peers));
                p = __shfl_sync(peers, b0, leader) + (u32)__popc(peers & lanemask_lt());
            }
        } else {
            if (act) p = atomicAdd(&mytab[src.digit(key[j])], 1u);
        }
        if (act) {
            if constexpr (Cfg::PACKED) {
                reinterpret_cast<u64*>(smem_raw)[p] = ((u64)src.out_key(key[j]) << 32) | (u64)val[j];
            } else {
                reinterpret_cast<Stage*>(smem_raw)[p] = key[j];
                pos[j] = p;
            }
            if constexpr (Cfg::HAS_AUX) saux[p] = aux[j];
        }
    }

    PSAC_PHASE(3);  // rank + scatter
    // ---- global offset of each bin of this tile: digit base + chunks before mine + tiles of my chunk before me
    if (tid < RADIX) goff[tid] = gbase[tid] + chunk_base[tid] + (u64)tile_excl[tid] - (u64)bin_start[tid];
    PSAC_PHASE(4);  // offsets
    __syncthreads();
    PSAC_PHASE(5);  // barrier 3

    // ---- coalesced write-out: consecutive shared positions of one bin are consecutive in global memory
    if constexpr (Cfg::PACKED) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int s = i * THREADS + tid;
            if (FULL || s < valid) {
                const u64 e = reinterpret_cast<const u64*>(smem_raw)[s];
                const Out k = (Out)(e >> 32);
                u8 a = 0;
                if constexpr (Cfg::HAS_AUX) a = saux[s];
                const u32 d = Src::FROM_TEXT ? (u32)a : src.digit((Stage)k);
                const u64 o = goff[d] + (u64)s;
                st_stream(kout + o, k);
                st_stream(vout + o, (ValT)(u32)e);
                if constexpr (Cfg::HAS_AUX) {
                    if (aout != nullptr) aout[o] = a;  // (pass 1 of the top-digit-first sort stages the digit only)
                }
            }
        }
    } else {
        const Stage* skeys = reinterpret_cast<const Stage*>(smem_raw);
        u8 dig[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int s = i * THREADS + tid;
            if (FULL || s < valid) {
                const Stage k = skeys[s];
                const u32 d = src.digit(k);
                dig[i] = (u8)d;
                const u64 o = goff[d] + (u64)s;
                if constexpr (Src::PEER)
                    src.kpeer[d][o] = src.out_key(k);  // bin d lives in (peer) GPU d's receive buffer
                else
                    st_stream(kout + o, src.out_key(k));
                if constexpr (Cfg::HAS_AUX) aout[o] = saux[s];
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
                if (FULL || s < valid) {
                    if constexpr (Src::PEER)
                        src.vpeer[dig[i]][goff[dig[i]] + (u64)s] = svals[s];
                    else
                        st_stream(vout + goff[dig[i]] + (u64)s, svals[s]);
                }
            }
        }
    }
    PSAC_PHASE(6);  // write-out
}

template <class Src, typename ValT, int THREADS, int ITEMS, bool HAS_AUX, int MINB, bool SAFE = false>
__global__ void __launch_bounds__(THREADS, MINB)
    radix_scatter_kernel(const Src src, typename Src::Out* __restrict__ kout, ValT* __restrict__ vout, u8* __restrict__ aout, size_t n,
                         const u64* __restrict__ gbase, const u64* __restrict__ chunk_base, const u32* __restrict__ tile_excl) {
    using Cfg = PassCfg<Src, ValT, THREADS, ITEMS, HAS_AUX>;
    constexpr int TILE = Cfg::TILE;
    static_assert(THREADS >= RADIX && THREADS % 32 == 0, "need one thread per digit");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* tab = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_TAB);
    for (int e = threadIdx.x; e < Cfg::NW * RADIX; e += THREADS) tab[e] = 0u;
    __syncthreads();
    const size_t tile = blockIdx.x;
    const size_t base = tile * (size_t)TILE;
    const u64* cb = chunk_base + (tile / SCAN_CHUNK) * RADIX;
    const u32* te = tile_excl + tile * RADIX;
    if (n - base >= (size_t)TILE)
        scatter_tile<Cfg, Src, ValT, true, SAFE>(smem_raw, src, kout, vout, aout, base, TILE, gbase, cb, te);
    else
        scatter_tile<Cfg, Src, ValT, false, SAFE>(smem_raw, src, kout, vout, aout, base, (int)(n - base), gbase, cb, te);
}

// per-tile digit histogram of a pass (same tile geometry as the scatter kernel): counts[tile][256]
template <class Src, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) tile_hist_kernel(const Src src, size_t n, u32* __restrict__ counts) {
    __shared__ u32 sh[RADIX];
    constexpr int TILE = THREADS * ITEMS;
    if (threadIdx.x < RADIX) sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * TILE;
    const int valid = (n - base >= (size_t)TILE) ? TILE : (int)(n - base);
    if constexpr (has_hist_digit<Src>::value) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int o = j * THREADS + threadIdx.x;
            if (o < valid) atomicAdd(&sh[src.hist_digit(base + o)], 1u);
        }
    } else {
        typename Src::Stage key[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int o = j * THREADS + threadIdx.x;
            if (o < valid) key[j] = src.load_key(base + o);
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (j * THREADS + (int)threadIdx.x < valid) atomicAdd(&sh[src.digit(key[j])], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < RADIX) counts[(size_t)blockIdx.x * RADIX + threadIdx.x] = sh[threadIdx.x];
}

// the same for digit pass 1 of a construction (keys from the packed text): a thread takes ITEMS consecutive suffixes and
// cuts their digits (the low `drop` key bits) out of a three-word window instead of loading two words per suffix
template <class Src, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) text_tile_hist_kernel(const Src src, size_t n, u32* __restrict__ counts) {
    __shared__ u32 sh[RADIX];
    constexpr int TILE = THREADS * ITEMS;
    if (threadIdx.x < RADIX) sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * TILE;
    const int valid = (n - base >= (size_t)TILE) ? TILE : (int)(n - base);
    const int o0 = (int)threadIdx.x * ITEMS;
    const int dbits = __popc(src.mask);  // bits of the digit
    if (base >= src.T && o0 + ITEMS <= valid && (ITEMS - 1) * src.lbits <= 64) {  // every digit starts inside the first two words
        const u64 bit0 = (base + o0 - src.T) * (u64)src.lbits + (u64)(src.kbits - src.dshift - dbits);  // first digit's stream position
        const u64 w0 = bit0 >> 6;
        const u64 a = __ldg(src.stream + w0), b = __ldg(src.stream + w0 + 1), c = __ldg(src.stream + w0 + 2);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const unsigned bit = (unsigned)(bit0 & 63) + (unsigned)(j * src.lbits);
            const u64 hi = bit < 64 ? a : b, lo = bit < 64 ? b : c;
            const unsigned o = bit & 63;
            const u64 v = o ? ((hi << o) | (lo >> (64 - o))) : hi;
            atomicAdd(&sh[(u32)(v >> (64 - dbits))], 1u);
        }
    } else {
        for (int j = 0; j < ITEMS; ++j)
            if (o0 + j < valid) atomicAdd(&sh[src.digit(src.load_key(base + o0 + j))], 1u);
    }
    __syncthreads();
    if (threadIdx.x < RADIX) counts[(size_t)blockIdx.x * RADIX + threadIdx.x] = sh[threadIdx.x];
}

// scan level 1: inside every chunk of SCAN_CHUNK tiles the counts become exclusive prefixes; chunk totals go out
__global__ void __launch_bounds__(RADIX) tile_scan_chunks_kernel(u32* __restrict__ counts, size_t tiles, u64* __restrict__ chunk_tot) {
    const size_t c = blockIdx.x, t0 = c * SCAN_CHUNK, t1 = (t0 + SCAN_CHUNK < tiles) ? t0 + SCAN_CHUNK : tiles;
    const int d = threadIdx.x;
    u32 run = 0;
#pragma unroll 8
    for (size_t t = t0; t < t1; ++t) {
        const u32 v = counts[t * RADIX + d];
        counts[t * RADIX + d] = run;
        run += v;
    }
    chunk_tot[c * RADIX + d] = run;
}

// scan level 2 (one CTA): chunk totals become exclusive prefixes over the chunks; digit totals -> exclusive digit bases.
// With seg_dense / seg_pad (top-digit-first suffix sort) the digit bases are also published as the 257-entry segment
// tables: dense starts, and starts padded to multiples of `pad_tile` elements -- gbase then holds the PADDED starts.
__global__ void __launch_bounds__(RADIX) tile_scan_top_kernel(u64* __restrict__ chunk_tot, size_t chunks, u64* __restrict__ gbase,
                                                              u64* __restrict__ seg_dense, u64* __restrict__ seg_pad, u64 pad_tile) {
    __shared__ u64 wtot[RADIX / 32], wtot2[RADIX / 32];
    const int d = threadIdx.x;
    u64 run = 0;
#pragma unroll 8
    for (size_t c = 0; c < chunks; ++c) {
        const u64 v = chunk_tot[c * RADIX + d];
        chunk_tot[c * RADIX + d] = run;
        run += v;
    }
    const u64 inc = warp_inclusive_scan(run, OpSum());
    const u64 padded = seg_pad != nullptr ? (run + pad_tile - 1) / pad_tile * pad_tile : 0;
    const u64 inc2 = warp_inclusive_scan(padded, OpSum());
    if ((d & 31) == 31) {
        wtot[d >> 5] = inc;
        wtot2[d >> 5] = inc2;
    }
    __syncthreads();
    u64 pre = 0, pre2 = 0;
    for (int w = 0; w < (d >> 5); ++w) {
        pre += wtot[w];
        pre2 += wtot2[w];
    }
    const u64 dense = pre + inc - run, pad = pre2 + inc2 - padded;
    if (gbase != nullptr) gbase[d] = seg_pad != nullptr ? pad : dense;
    if (seg_pad != nullptr) {
        seg_dense[d] = dense;
        seg_pad[d] = pad;
        if (d == RADIX - 1) {
            seg_dense[RADIX] = dense + run;
            seg_pad[RADIX] = pad + padded;
        }
    }
}

// ------------------------------------------------------------------ top-digit-first sort of suffixes: segmented passes
// After the first pass has partitioned the suffixes by the TOP digit of their key (256 segments), the remaining digits
// are sorted LSD *inside each segment*, so the top digit never has to be stored: it is implied by the segment.  Segments
// start on multiples of the tile size in the ping-pong buffers ("padded" layout), so a tile never straddles two
// segments; the last pass writes the dense layout.  tile_info[t] = segment << 16 | valid elements of padded tile t.
template <int TILE>
__global__ void __launch_bounds__(256) seg_tiles_kernel(const u64* __restrict__ seg_dense, const u64* __restrict__ seg_pad, u32* __restrict__ tile_info,
                                                        size_t rows) {
    __shared__ u64 s_pad[RADIX + 1];
    for (int e = threadIdx.x; e <= RADIX; e += blockDim.x) s_pad[e] = seg_pad[e];
    __syncthreads();
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows; t += (size_t)gridDim.x * blockDim.x) {
        const u64 pos = (u64)t * TILE;
        int lo = 0, hi = RADIX;  // last segment whose padded start is <= pos
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_pad[mid] <= pos)
                lo = mid;
            else
                hi = mid;
        }
        // empty segments share their start with the next one: move to the last segment starting here
        while (lo + 1 < RADIX && s_pad[lo + 1] <= pos) ++lo;
        const u64 len = seg_dense[lo + 1] - seg_dense[lo];
        const u64 off = pos - s_pad[lo];
        const u32 valid = off < len ? (u32)((len - off < (u64)TILE) ? len - off : (u64)TILE) : 0u;
        tile_info[t] = ((u32)lo << 16) | valid;
    }
}

template <class Src, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) tile_hist_seg_kernel(const Src src, const u32* __restrict__ tile_info, u32* __restrict__ counts) {
    __shared__ u32 sh[RADIX];
    constexpr int TILE = THREADS * ITEMS;
    if (threadIdx.x < RADIX) sh[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * TILE;
    const int valid = (int)(tile_info[blockIdx.x] & 0xffffu);
    typename Src::Stage key[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int o = j * THREADS + threadIdx.x;
        if (o < valid) key[j] = src.load_key(base + o);
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        if (j * THREADS + (int)threadIdx.x < valid) atomicAdd(&sh[src.digit(key[j])], 1u);
    }
    __syncthreads();
    if (threadIdx.x < RADIX) counts[(size_t)blockIdx.x * RADIX + threadIdx.x] = sh[threadIdx.x];
}

// per segment: digit bases of this pass.  P(t, d) = number of digit-d elements in the padded tiles before t (from the
// two-level scan).  Segment s covers tiles [a, b): segbase[s][d] = out_start[s] + sum_{d' < d} (P(b,d') - P(a,d')) - P(a,d),
// so that an element's output position is segbase[s][d] + P(t, d) + its rank inside the tile.
template <int TILE>
__global__ void __launch_bounds__(RADIX) seg_base_kernel(const u32* __restrict__ tile_excl, const u64* __restrict__ chunk_base,
                                                         const u64* __restrict__ seg_dense, const u64* __restrict__ seg_pad, int dense_out,
                                                         u64* __restrict__ segbase) {
    __shared__ u64 wtot[RADIX / 32];
    const int s = blockIdx.x, d = threadIdx.x;
    const u64 len = seg_dense[s + 1] - seg_dense[s];
    const size_t a = (size_t)(seg_pad[s] / TILE), b = a + (size_t)((len + TILE - 1) / TILE);
    const u64 pa = chunk_base[(a / SCAN_CHUNK) * RADIX + d] + (u64)tile_excl[a * RADIX + d];
    const u64 pb = chunk_base[(b / SCAN_CHUNK) * RADIX + d] + (u64)tile_excl[b * RADIX + d];
    const u64 tot = pb - pa;
    const u64 inc = warp_inclusive_scan(tot, OpSum());
    if ((d & 31) == 31) wtot[d >> 5] = inc;
    __syncthreads();
    u64 pre = 0;
    for (int w = 0; w < (d >> 5); ++w) pre += wtot[w];
    segbase[(size_t)s * RADIX + d] = (dense_out ? seg_dense[s] : seg_pad[s]) + (pre + inc - tot) - pa;
}

template <class Src, typename ValT, int THREADS, int ITEMS, int MINB, bool SAFE = false>
__global__ void __launch_bounds__(THREADS, MINB)
    radix_scatter_seg_kernel(const Src src, typename Src::Out* __restrict__ kout, ValT* __restrict__ vout, const u32* __restrict__ tile_info,
                             const u64* __restrict__ segbase, const u64* __restrict__ chunk_base, const u32* __restrict__ tile_excl) {
    using Cfg = PassCfg<Src, ValT, THREADS, ITEMS, false>;
    constexpr int TILE = Cfg::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 info = tile_info[blockIdx.x];
    const int valid = (int)(info & 0xffffu);
    if (valid == 0) return;
    u32* tab = reinterpret_cast<u32*>(smem_raw + Cfg::OFF_TAB);
    for (int e = threadIdx.x; e < Cfg::NW * RADIX; e += THREADS) tab[e] = 0u;
    __syncthreads();
    const size_t tile = blockIdx.x;
    const size_t base = tile * (size_t)TILE;
    const u64* gb = segbase + (size_t)(info >> 16) * RADIX;
    const u64* cb = chunk_base + (tile / SCAN_CHUNK) * RADIX;
    const u32* te = tile_excl + tile * RADIX;
    if (valid == TILE)
        scatter_tile<Cfg, Src, ValT, true, SAFE>(smem_raw, src, kout, vout, nullptr, base, TILE, gb, cb, te);
    else
        scatter_tile<Cfg, Src, ValT, false, SAFE>(smem_raw, src, kout, vout, nullptr, base, valid, gb, cb, te);
}

// ------------------------------------------------------------------ hardware self-test of the ranking assumption
// Replays the ranking loop of scatter_tile (ITEMS back-to-back ATOMS.ADD per lane on a per-warp table, some lanes
// inactive) and compares every returned value with the stable rank computed from ballots.  Returns the number of
// mismatches in *bad; the engine refuses to run when it is not zero.
template <int ITEMS>
__global__ void __launch_bounds__(512) atoms_order_selftest_kernel(u32 seed, u32 ndig, unsigned long long* bad) {
    __shared__ u32 tab[16][RADIX];
    __shared__ u32 ref[16][RADIX];
    const u32 warp = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    unsigned long long mism = 0;
    u32 x = seed * 2654435761u + (blockIdx.x * blockDim.x + threadIdx.x) * 40503u + 977u;
    for (int round = 0; round < 8; ++round) {
        for (int e = threadIdx.x; e < 16 * RADIX; e += blockDim.x) {
            (&tab[0][0])[e] = 0;
            (&ref[0][0])[e] = 0;
        }
        __syncthreads();
        u32 d[ITEMS], got[ITEMS];
        bool act[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            x = x * 1664525u + 1013904223u;
            d[j] = (x >> 16) % ndig;
            act[j] = ((x >> 8) & 7u) != 0u || (round & 1);
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            got[j] = 0;
            if (act[j]) got[j] = atomicAdd(&tab[warp][d[j]], 1u);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u32 peers = __ballot_sync(0xffffffffu, act[j]);
#pragma unroll
            for (int b = 0; b < RADIX_BITS; ++b) {
                const u32 m = __ballot_sync(0xffffffffu, (d[j] >> b) & 1u);
                peers &= ((d[j] >> b) & 1u) ? m : ~m;
            }
            const u32 before = ref[warp][d[j]];
            __syncwarp();
            if (act[j]) {
                mism += (got[j] != before + __popc(peers & lt)) ? 1u : 0u;
                if ((peers & lt) == 0) ref[warp][d[j]] = before + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
    }
    if (mism) atomicAdd(bad, mism);
}

// ------------------------------------------------------------------ host driver
// Tile shape of a pass (swept with tools/bench_pass.cu, profiles/r1_bench_pass_sweep.txt): 32-bit pairs run best with
// 512 threads and 2 CTAs per SM -- 12 keys per thread when the auxiliary byte travels along, 16 without it; wider pairs
// keep 384 threads.
template <typename OutT, typename ValT, bool HAS_AUX>
struct SortTuning {
    static constexpr bool NARROW = (sizeof(OutT) == 4 && sizeof(ValT) <= 4) || (sizeof(OutT) == 8 && std::is_same<ValT, NoVal>::value);  // 8 staged bytes
    static constexpr int THREADS = NARROW ? 512 : 384;
    static constexpr int ITEMS = NARROW ? (HAS_AUX ? 12 : 16) : ((sizeof(OutT) + sizeof(ValT) >= 16) ? 12 : 16);
    static constexpr int MINB = 2;
};
constexpr int MIN_TILE = 256 * 8;  // smallest tile any configuration uses (sizes the per-tile workspace)

// function attributes are per device: true the first time a call site runs on the current device
static inline bool first_use_on_device(bool (&seen)[64]) {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (seen[d]) return false;
    seen[d] = true;
    return true;
}

// Ranking mode of every digit pass launched from this process: false = one ATOMS.ADD per key (hardware lane order,
// verified by atoms_order_selftest_kernel), true = match.any ranking.  Set at engine creation.
static bool g_safe_rank = false;

// launches `KERN<..., false>` or `KERN<..., true>` according to g_safe_rank; both get the dynamic shared memory attribute
#define PSAC_LAUNCH_RANKED(KERN_FAST, KERN_SAFE, SMEM, GRID, THREADS, STREAM, ...)                                              \
    do {                                                                                                                        \
        static bool _seen[64] = {};                                                                                             \
        if (first_use_on_device(_seen)) {                                                                                       \
            PSAC_CUDA(cudaFuncSetAttribute(KERN_FAST, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));               \
            PSAC_CUDA(cudaFuncSetAttribute(KERN_SAFE, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));               \
        }                                                                                                                       \
        if (g_safe_rank)                                                                                                        \
            KERN_SAFE<<<(GRID), (THREADS), (SMEM), (STREAM)>>>(__VA_ARGS__);                                                    \
        else                                                                                                                    \
            KERN_FAST<<<(GRID), (THREADS), (SMEM), (STREAM)>>>(__VA_ARGS__);                                                    \
    } while (0)

struct RadixWorkspace {
    u64* gbase = nullptr;      // [RADIX] digit bases of the pass in flight
    void* tiles = nullptr;     // per-tile counts (u32 [tiles][RADIX]) followed by the chunk totals (u64 [chunks][RADIX])
    size_t tiles_bytes = 0;
    static size_t tiles_bytes_for(size_t n) { return (div_up(n ? n : 1, (size_t)MIN_TILE) + 2) * RADIX * sizeof(u64); }
};

// One digit pass: histogram, two-level scan, scatter.  4 launches.  (THREADS, ITEMS, MINB = CTAs per SM the scatter kernel
// is compiled for; tools/bench_pass.cu sweeps them, SortTuning holds the choice.)
template <class Src, typename ValT, bool HAS_AUX, int THREADS, int ITEMS, int MINB>
void launch_pass_cfg(const RadixWorkspace& ws, const Src& src, typename Src::Out* kout, ValT* vout, u8* aout, size_t n, cudaStream_t stream) {
    using Cfg = PassCfg<Src, ValT, THREADS, ITEMS, HAS_AUX>;
    const size_t tiles = div_up(n, (size_t)Cfg::TILE);
    const size_t chunks = div_up(tiles, (size_t)SCAN_CHUNK);
    const size_t counts_bytes = align_up(tiles * RADIX * sizeof(u32), 256);
    if (counts_bytes + chunks * RADIX * sizeof(u64) > ws.tiles_bytes) throw std::string("radix pass: tile workspace too small");
    u32* counts = reinterpret_cast<u32*>(ws.tiles);
    u64* chunk_tot = reinterpret_cast<u64*>(reinterpret_cast<char*>(ws.tiles) + counts_bytes);
    auto kern = radix_scatter_kernel<Src, ValT, THREADS, ITEMS, HAS_AUX, MINB, false>;
    auto kern_safe = radix_scatter_kernel<Src, ValT, THREADS, ITEMS, HAS_AUX, MINB, true>;
    if constexpr (Src::FROM_TEXT)
        text_tile_hist_kernel<Src, THREADS, ITEMS><<<(unsigned)tiles, THREADS, 0, stream>>>(src, n, counts);
    else
        tile_hist_kernel<Src, THREADS, ITEMS><<<(unsigned)tiles, THREADS, 0, stream>>>(src, n, counts);
    tile_scan_chunks_kernel<<<(unsigned)chunks, RADIX, 0, stream>>>(counts, tiles, chunk_tot);
    tile_scan_top_kernel<<<1, RADIX, 0, stream>>>(chunk_tot, chunks, ws.gbase, nullptr, nullptr, 1);
    PSAC_LAUNCH_RANKED(kern, kern_safe, Cfg::SMEM, (unsigned)tiles, THREADS, stream, src, kout, vout, aout, n, ws.gbase, chunk_tot, counts);
}

template <class Src, typename ValT, bool HAS_AUX>
void launch_pass(const RadixWorkspace& ws, const Src& src, typename Src::Out* kout, ValT* vout, u8* aout, size_t n, cudaStream_t stream) {
    using T = SortTuning<typename Src::Out, ValT, HAS_AUX>;
    launch_pass_cfg<Src, ValT, HAS_AUX, T::THREADS, T::ITEMS, T::MINB>(ws, src, kout, vout, aout, n, stream);
}
constexpr int LAUNCHES_PER_PASS = 4;

// Sorts n pairs by key bits [begin_bit, end_bit).  Ping-pongs between (keys, vals) and (keys_alt, vals_alt);
// returns true when the sorted data ended up in the *_alt buffers.
template <typename KeyT, typename ValT>
bool radix_sort_pairs(const RadixWorkspace& ws, KeyT* keys, KeyT* keys_alt, ValT* vals, ValT* vals_alt, size_t n, int begin_bit, int end_bit,
                      cudaStream_t stream, int sm_count, RadixPlan* plan_out = nullptr, uint64_t* launches = nullptr) {
    (void)sm_count;
    RadixPlan plan = make_radix_plan(begin_bit, end_bit);
    if (plan_out) *plan_out = plan;
    if (n == 0 || plan.npass == 0) return false;
    if (launches) *launches = (uint64_t)LAUNCHES_PER_PASS * plan.npass;
    bool in_alt = false;
    for (int p = 0; p < plan.npass; ++p) {
        ArraySrc<KeyT, ValT> src{in_alt ? keys_alt : keys, in_alt ? vals_alt : vals, nullptr, plan.shift[p], (1u << plan.bits[p]) - 1u, (KeyT)0};
        launch_pass<ArraySrc<KeyT, ValT>, ValT, false>(ws, src, in_alt ? keys : keys_alt, in_alt ? vals : vals_alt, nullptr, n, stream);
        in_alt = !in_alt;
    }
    PSAC_CUDA(cudaGetLastError());
    return in_alt;
}

// ---------------------------------------------------------------------------------------------------------------
// First sort of a construction, 64-bit carried keys (more than 40 key bits, or the |Sigma| = 256 quirk): plain LSD, the
// first pass reads the packed text and carries the whole key.  kbuf / vbuf are two ping-pong buffers; pass 1 writes
// buffer 0; returns the index (0/1) of the buffers holding the sorted keys and suffix indices.
template <typename IdxT>
int radix_sort_suffixes(const RadixWorkspace& ws, const u64* text_stream, size_t n, int lbits, int key_chars, u64* const kbuf[2], IdxT* const vbuf[2],
                        cudaStream_t stream, RadixPlan* plan_out, uint64_t* launches, cudaEvent_t ev_pass1_done = nullptr) {
    const int kbits = key_chars * lbits;
    RadixPlan plan = make_radix_plan(0, kbits);
    if (plan_out) *plan_out = plan;
    if (n == 0) return 0;
    if (launches) *launches = (uint64_t)LAUNCHES_PER_PASS * plan.npass;
    {
        const u64 C = (u64)key_chars;
        TextSrc<u64, IdxT> src{text_stream, (u64)n, (n < C - 1) ? (u64)n : C - 1, lbits, kbits, 0, (1u << plan.bits[0]) - 1u, 0};
        launch_pass<TextSrc<u64, IdxT>, IdxT, false>(ws, src, kbuf[0], vbuf[0], nullptr, n, stream);
    }
    if (ev_pass1_done) cudaEventRecord(ev_pass1_done, stream);
    int cur = 0;
    for (int p = 1; p < plan.npass; ++p) {
        ArraySrc<u64, IdxT> src{kbuf[cur], vbuf[cur], nullptr, plan.shift[p], (1u << plan.bits[p]) - 1u, 0ull};
        launch_pass<ArraySrc<u64, IdxT>, IdxT, false>(ws, src, kbuf[1 - cur], vbuf[1 - cur], nullptr, n, stream);
        cur = 1 - cur;
    }
    PSAC_CUDA(cudaGetLastError());
    return cur;
}

// First sort of a construction, 32-bit carried keys (key bits minus the top digit <= 32: BASELINE configs[1]).
// Top digit first: pass 1 reads the packed text and partitions the suffixes by the TOP digit of their key into 256
// segments laid out on tile boundaries; the remaining digits are then sorted LSD inside the segments
// (tile_hist_seg_kernel / seg_base_kernel / radix_scatter_seg_kernel), the last pass writing the dense layout.  The
// keys carried between passes are the LOW bits only -- the top digit is implied by the segment (seg_dense[257] = dense
// segment starts, returned for the resolve step), so a suffix moves as 4 + 4 bytes per pass and no byte array travels.
// kbuf / vbuf: two ping-pong buffers of suffix_sort_padded_elems(n) elements; vfinal: n elements (may be the caller's
// SA buffer).  Returns the index of the kbuf holding the sorted carried keys (dense).
static inline size_t suffix_sort_tile() { return (size_t)SortTuning<u32, u32, false>::THREADS * SortTuning<u32, u32, false>::ITEMS; }
static inline size_t suffix_sort_padded_elems(size_t n) { return n + (size_t)(RADIX + 1) * suffix_sort_tile(); }
static inline size_t suffix_sort_rows(size_t n) { return div_up(suffix_sort_padded_elems(n), suffix_sort_tile()) + 1; }

struct SegWorkspace {
    u64* seg_dense;  // [257]
    u64* seg_pad;    // [257]
    u64* segbase;    // [256][256]
    u32* tile_info;  // [suffix_sort_rows(n)]
};

template <typename IdxT>
int radix_sort_suffixes_msd(const RadixWorkspace& ws, const SegWorkspace& sw, const u64* text_stream, size_t n, int lbits, int key_chars, u32* const kbuf[2],
                            IdxT* const vbuf[2], IdxT* vfinal, cudaStream_t stream, RadixPlan* plan_out, uint64_t* launches,
                            cudaEvent_t ev_pass1_done = nullptr, cudaEvent_t* ev_scatter = nullptr) {
    // ev_scatter (optional, 2 * MAX_PASSES events): [2p] / [2p + 1] bracket the scatter kernel of segmented pass p
    using T = SortTuning<u32, IdxT, false>;
    constexpr int TILE = T::THREADS * T::ITEMS;
    const int kbits = key_chars * lbits;
    RadixPlan plan = make_radix_plan(0, kbits);
    if (plan_out) *plan_out = plan;
    if (n == 0) return 0;
    const int top = plan.npass - 1;
    if (kbits - plan.bits[top] > 32) throw std::string("radix_sort_suffixes_msd: carried key does not fit 32 bits");
    const u64 C = (u64)key_chars;
    const u64 T0 = (n < C - 1) ? (u64)n : C - 1;
    uint64_t nl = 0;
    // ---- pass 1: top digit, from the text.  A single-pass sort (<= 8 key bits) writes the dense layout at once.
    {
        // (the digit is staged as the aux byte in shared memory -- the packed pair no longer contains it -- but not written out)
        using TT = SortTuning<u32, IdxT, true>;
        using Src = TextSrc<u32, IdxT>;
        using Cfg = PassCfg<Src, IdxT, TT::THREADS, TT::ITEMS, true>;
        Src src{text_stream, (u64)n, T0, lbits, kbits, 0, (1u << plan.bits[top]) - 1u, plan.shift[top]};
        const size_t tiles = div_up(n, (size_t)Cfg::TILE);
        const size_t chunks = div_up(tiles, (size_t)SCAN_CHUNK);
        const size_t counts_bytes = align_up(tiles * RADIX * sizeof(u32), 256);
        if (counts_bytes + chunks * RADIX * sizeof(u64) > ws.tiles_bytes) throw std::string("radix pass: tile workspace too small");
        u32* counts = reinterpret_cast<u32*>(ws.tiles);
        u64* chunk_tot = reinterpret_cast<u64*>(reinterpret_cast<char*>(ws.tiles) + counts_bytes);
        auto kern = radix_scatter_kernel<Src, IdxT, TT::THREADS, TT::ITEMS, true, TT::MINB, false>;
        auto kern_safe = radix_scatter_kernel<Src, IdxT, TT::THREADS, TT::ITEMS, true, TT::MINB, true>;
        const bool single = plan.npass == 1;
        text_tile_hist_kernel<Src, TT::THREADS, TT::ITEMS><<<(unsigned)tiles, TT::THREADS, 0, stream>>>(src, n, counts);
        tile_scan_chunks_kernel<<<(unsigned)chunks, RADIX, 0, stream>>>(counts, tiles, chunk_tot);
        // gbase = padded segment starts (dense ones for a single-pass sort: pad_tile = 1)
        tile_scan_top_kernel<<<1, RADIX, 0, stream>>>(chunk_tot, chunks, ws.gbase, sw.seg_dense, sw.seg_pad, single ? 1 : (u64)TILE);
        PSAC_LAUNCH_RANKED(kern, kern_safe, Cfg::SMEM, (unsigned)tiles, TT::THREADS, stream, src, kbuf[0], single ? vfinal : vbuf[0], nullptr, n, ws.gbase,
                           chunk_tot, counts);
        nl += 4;
        if (ev_pass1_done) cudaEventRecord(ev_pass1_done, stream);
        if (single) {
            if (launches) *launches = nl;
            PSAC_CUDA(cudaGetLastError());
            return 0;
        }
    }
    // ---- tile map of the padded layout
    const size_t rows = suffix_sort_rows(n);
    seg_tiles_kernel<TILE><<<(unsigned)div_up(rows, (size_t)256), 256, 0, stream>>>(sw.seg_dense, sw.seg_pad, sw.tile_info, rows);
    nl += 1;
    // ---- the remaining digits, least significant first, inside the segments
    using Src = ArraySrc<u32, IdxT>;
    using Cfg = PassCfg<Src, IdxT, T::THREADS, T::ITEMS, false>;
    auto kern = radix_scatter_seg_kernel<Src, IdxT, T::THREADS, T::ITEMS, T::MINB, false>;
    auto kern_safe = radix_scatter_seg_kernel<Src, IdxT, T::THREADS, T::ITEMS, T::MINB, true>;
    const size_t chunks = div_up(rows, (size_t)SCAN_CHUNK);
    const size_t counts_bytes = align_up(rows * RADIX * sizeof(u32), 256);
    if (counts_bytes + chunks * RADIX * sizeof(u64) > ws.tiles_bytes) throw std::string("radix pass: tile workspace too small");
    u32* counts = reinterpret_cast<u32*>(ws.tiles);
    u64* chunk_tot = reinterpret_cast<u64*>(reinterpret_cast<char*>(ws.tiles) + counts_bytes);
    int cur = 0;
    for (int p = 0; p < top; ++p) {
        const bool last = p == top - 1;
        Src src{kbuf[cur], vbuf[cur], nullptr, plan.shift[p], (1u << plan.bits[p]) - 1u, 0u};
        tile_hist_seg_kernel<Src, T::THREADS, T::ITEMS><<<(unsigned)rows, T::THREADS, 0, stream>>>(src, sw.tile_info, counts);
        tile_scan_chunks_kernel<<<(unsigned)chunks, RADIX, 0, stream>>>(counts, rows, chunk_tot);
        tile_scan_top_kernel<<<1, RADIX, 0, stream>>>(chunk_tot, chunks, nullptr, nullptr, nullptr, 1);
        seg_base_kernel<TILE><<<RADIX, RADIX, 0, stream>>>(counts, chunk_tot, sw.seg_dense, sw.seg_pad, last ? 1 : 0, sw.segbase);
        if (ev_scatter) cudaEventRecord(ev_scatter[2 * p], stream);
        PSAC_LAUNCH_RANKED(kern, kern_safe, Cfg::SMEM, (unsigned)rows, T::THREADS, stream, src, kbuf[1 - cur], last ? vfinal : vbuf[1 - cur], sw.tile_info,
                           sw.segbase, chunk_tot, counts);
        if (ev_scatter) cudaEventRecord(ev_scatter[2 * p + 1], stream);
        nl += 5;
        cur = 1 - cur;
    }
    if (launches) *launches = nl;
    PSAC_CUDA(cudaGetLastError());
    return cur;
}

// ---------------------------------------------------------------------------------------------------------------
// Sharded construction (sharded.cuh): the LSD digit passes over 64-bit words [carried key | suffix index] inside the
// top-digit segments a rank received from its peers (TextWordSrc).  sw.seg_dense / sw.seg_pad are given by the caller
// (257 entries each, from the all-gathered digit counts); W[0] holds the padded layout, `rows` padded tiles.  Key bits
// [bit0, bit0 + nbits) of the words are sorted; the last pass writes the dense layout.  Returns the index of the buffer
// holding the dense result.
static inline size_t word_sort_tile() { return (size_t)SortTuning<u64, NoVal, false>::THREADS * SortTuning<u64, NoVal, false>::ITEMS; }

static inline int radix_sort_words_seg(const RadixWorkspace& ws, const SegWorkspace& sw, size_t rows, u64* const W[2], int bit0, int nbits, cudaStream_t stream,
                                       RadixPlan* plan_out, uint64_t* launches, cudaEvent_t* ev_scatter = nullptr) {
    using T = SortTuning<u64, NoVal, false>;
    constexpr int TILE = T::THREADS * T::ITEMS;
    using Src = ArraySrc<u64, NoVal>;
    using Cfg = PassCfg<Src, NoVal, T::THREADS, T::ITEMS, false>;
    RadixPlan plan = make_radix_plan(bit0, bit0 + nbits);
    if (plan_out) *plan_out = plan;
    uint64_t nl = 0;
    seg_tiles_kernel<TILE><<<(unsigned)div_up(rows, (size_t)256), 256, 0, stream>>>(sw.seg_dense, sw.seg_pad, sw.tile_info, rows);
    nl += 1;
    auto kern = radix_scatter_seg_kernel<Src, NoVal, T::THREADS, T::ITEMS, T::MINB, false>;
    auto kern_safe = radix_scatter_seg_kernel<Src, NoVal, T::THREADS, T::ITEMS, T::MINB, true>;
    const size_t chunks = div_up(rows, (size_t)SCAN_CHUNK);
    const size_t counts_bytes = align_up(rows * RADIX * sizeof(u32), 256);
    if (counts_bytes + chunks * RADIX * sizeof(u64) > ws.tiles_bytes) throw std::string("radix pass: tile workspace too small");
    u32* counts = reinterpret_cast<u32*>(ws.tiles);
    u64* chunk_tot = reinterpret_cast<u64*>(reinterpret_cast<char*>(ws.tiles) + counts_bytes);
    int cur = 0;
    for (int p = 0; p < plan.npass; ++p) {
        const bool last = p == plan.npass - 1;
        Src src{W[cur], nullptr, nullptr, plan.shift[p], (1u << plan.bits[p]) - 1u, 0ull};
        tile_hist_seg_kernel<Src, T::THREADS, T::ITEMS><<<(unsigned)rows, T::THREADS, 0, stream>>>(src, sw.tile_info, counts);
        tile_scan_chunks_kernel<<<(unsigned)chunks, RADIX, 0, stream>>>(counts, rows, chunk_tot);
        tile_scan_top_kernel<<<1, RADIX, 0, stream>>>(chunk_tot, chunks, nullptr, nullptr, nullptr, 1);
        seg_base_kernel<TILE><<<RADIX, RADIX, 0, stream>>>(counts, chunk_tot, sw.seg_dense, sw.seg_pad, last ? 1 : 0, sw.segbase);
        if (ev_scatter) cudaEventRecord(ev_scatter[2 * p], stream);
        PSAC_LAUNCH_RANKED(kern, kern_safe, Cfg::SMEM, (unsigned)rows, T::THREADS, stream, src, W[1 - cur], (NoVal*)nullptr, sw.tile_info, sw.segbase, chunk_tot,
                           counts);
        if (ev_scatter) cudaEventRecord(ev_scatter[2 * p + 1], stream);
        nl += 5;
        cur = 1 - cur;
    }
    if (launches) *launches = nl;
    PSAC_CUDA(cudaGetLastError());
    return cur;
}

}  // namespace psacb200
