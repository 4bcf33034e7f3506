// psac-b200: shared device/host helpers for the sm_100a suffix-array engine.
//
// Everything on this path is unsigned-integer indexing work bounded by HBM bandwidth
// (SURVEY.md section 8d): no tensor cores, no floating point.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

namespace psacb200 {

// ---------------------------------------------------------------- error plumbing
void set_last_error(const std::string& msg);

struct cuda_failure {
    cudaError_t err;
    const char* what;
    const char* file;
    int line;
};

#define PSAC_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) throw psacb200::cuda_failure{_e, #call, __FILE__, __LINE__}; \
    } while (0)

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 64-bit entries shared between CTAs for the decoupled look-back: one aligned 8-byte word carries
// payload + epoch + state, so a single relaxed load observes a consistent triple.
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming (evict-first) accessors for arrays that are touched exactly once per kernel
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
    return __ldcs(p);
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) {
    __stcs(p, v);
}

__host__ __device__ __forceinline__ unsigned bits_for(u64 x) {  // number of bits needed to store x (0 -> 0)
    unsigned b = 0;
    while (x) {
        ++b;
        x >>= 1;
    }
    return b;
}

static inline size_t div_up(size_t a, size_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return div_up(a, b) * b; }

// ---------------------------------------------------------------- packed text
// The text is stored as a big-endian bit stream of dense codes, lbits in {1,2,4,8} per character, character i
// at stream bits [i*lbits, (i+1)*lbits).  Word w holds characters [w*cpw, (w+1)*cpw), first character in the
// most significant bits, zero past the end; the stream has two zero words of padding.
struct CodeTable {
    u8 code[256];
};

// nbits (1..64) stream bits starting at stream bit position `bit`, right-aligned
__device__ __forceinline__ u64 stream_bits(const u64* __restrict__ stream, u64 bit, int nbits) {
    const u64 w = bit >> 6;
    const unsigned o = (unsigned)(bit & 63);
    const u64 hi = __ldg(stream + w), lo = __ldg(stream + w + 1);
    const u64 v = o ? ((hi << o) | (lo >> (64 - o))) : hi;
    return v >> (64 - nbits);
}
// nbits (<= 64) stream bits starting at character position i, right-aligned
__device__ __forceinline__ u64 stream_extract(const u64* __restrict__ stream, u64 i, int lbits, int nbits) {
    return stream_bits(stream, i * (u64)lbits, nbits);
}

// ---------------------------------------------------------------- generalized suffix array (string sets)
// In a string-set construction code 0 is the separator: it ends its string, and no comparison runs past it
// (reference kmer_gen_stringset, include/kmer.hpp:269-355: k-mers are 0-filled behind the end of their string).
// lowest bit of every lbits-wide field of a word (lbits in {1,2,4,8})
__host__ __device__ __forceinline__ u64 field_lsb(int lbits) {
    return lbits == 1 ? ~0ull : lbits == 2 ? 0x5555555555555555ull : lbits == 4 ? 0x1111111111111111ull : 0x0101010101010101ull;
}
// the lowest bit of a field is set in the result iff the field of w is zero
__host__ __device__ __forceinline__ u64 zero_fields(u64 w, int lbits) {
    u64 nz = w;
    if (lbits >= 2) nz |= nz >> 1;
    if (lbits >= 4) nz |= nz >> 2;
    if (lbits == 8) nz |= nz >> 4;
    return ~nz & field_lsb(lbits);
}
// key of kbits bits (whole characters, right-aligned): everything behind its first zero character is cleared
__device__ __forceinline__ u64 gsa_mask_key(u64 k, int lbits, int kbits) {
    u64 z = zero_fields(k, lbits);
    if (kbits < 64) z &= (1ull << kbits) - 1ull;
    if (z == 0) return k;
    const int p = 63 - __clzll((long long)z);  // lowest bit of the first zero character
    return k & ~((1ull << p) - 1ull);
}

// ---------------------------------------------------------------- block distribution on the device
// mxx::blk_dist (reference ext/mxx/include/mxx/partition.hpp:283-331): n elements over p ranks, the first n % p ranks
// hold one element more.  owner() estimates the quotient with one 32-bit multiply-high and corrects it against the block starts.
struct BlkDiv {
    u64 cut, base1, base;  // cut = rem * (base + 1); base1 = base + 1
    u32 rem, p;
    u32 sh, magic32;       // quotient estimate: umulhi(g >> sh, magic32) <= g / block size, at most 2 below it
    __host__ static BlkDiv make(u64 n, int p) {
        BlkDiv d;
        const u64 b = n / (u64)p;
        d.p = (u32)p;
        d.rem = (u32)(n % (u64)p);
        d.base = b;
        d.base1 = b + 1;
        d.cut = (u64)d.rem * (b + 1);
        d.sh = 0;
        while ((n >> d.sh) >> 32) ++d.sh;
        const u64 d32 = (d.base1 >> d.sh) + 1;  // over-estimates the divisor: the quotient estimate never overshoots
        d.magic32 = d32 >= 2 ? (u32)((1ull << 32) / d32) : 0xffffffffu;
        return d;
    }
    __device__ __forceinline__ u64 start(u32 r) const { return (u64)r * base + (u64)(r < rem ? r : rem); }
    // owner of global index g (g < n) and its index inside the owner's block; integer instructions only
    __device__ __forceinline__ u32 owner(u64 g, u64* local) const {
        u32 q = __umulhi((u32)(g >> sh), magic32);
        q = q < p - 1 ? q : p - 1;
        u64 st = start(q);
        if (q + 1 < p) {
            u64 nx = start(q + 1);
            if (g >= nx) {
                ++q;
                st = nx;
                if (q + 1 < p) {
                    nx = start(q + 1);
                    if (g >= nx) {
                        ++q;
                        st = nx;
                        // (blocks smaller than 2^16 elements: the estimate may be further off)
                        while (q + 1 < p && g >= start(q + 1)) st = start(++q);
                    }
                }
            }
        }
        *local = g - st;
        return q;
    }
};

// The suffix-index field of a sort word of the sharded construction: (owning rank of the suffix's ISA entry, index inside
// that rank's text block) instead of the global index -- every later step that routes by owner (the SA -> ISA exchange)
// reads it with two shifts, and the owner is computed once per 64 characters where the word is made.
struct WordIdx {
    BlkDiv div;
    int lb;    // bits of the block-local index
    u64 mask;  // mask of the whole field (lb + rank bits)
    __device__ __forceinline__ u64 decode(u64 w) const {
        const u64 f = w & mask;
        return div.start((u32)(f >> lb)) + (f & ((1ull << lb) - 1ull));
    }
    __device__ __forceinline__ u64 encode(u64 g) const {
        u64 local;
        const u64 o = div.owner(g, &local);
        return (o << lb) | local;
    }
    __device__ __forceinline__ u64 block_size(u32 r) const { return div.base + (r < div.rem ? 1u : 0u); }
};

// ---------------------------------------------------------------- decoupled look-back channel
// One u64 per tile: [63:10] payload (54 bits), [9:2] epoch, [1:0] state.  The epoch makes entries of a
// previous use of the same buffer read as "not yet written" without a memset between uses.
enum : u64 { LB_NONE = 0, LB_AGGREGATE = 1, LB_INCLUSIVE = 2 };

__device__ __forceinline__ u64 lb_pack(u64 payload, u32 epoch, u64 state) { return (payload << 10) | ((u64)(epoch & 0xffu) << 2) | state; }
__device__ __forceinline__ u64 lb_payload(u64 w) { return w >> 10; }
__device__ __forceinline__ u64 lb_state(u64 w, u32 epoch) { return (((w >> 2) & 0xffu) == (epoch & 0xffu)) ? (w & 3u) : (u64)LB_NONE; }

struct OpSum {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a + b; }
    static __device__ __forceinline__ u64 identity() { return 0; }
};
struct OpMax {
    __device__ __forceinline__ u64 operator()(u64 a, u64 b) const { return a > b ? a : b; }
    static __device__ __forceinline__ u64 identity() { return 0; }
};

// ---------------------------------------------------------------- warp scans (shuffle based)
template <typename Op>
__device__ __forceinline__ u64 warp_inclusive_scan(u64 v, Op op) {
    unsigned lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v = op(o, v);
    }
    return v;
}
__device__ __forceinline__ u32 warp_inclusive_sum_u32(u32 v) {
    unsigned lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// Look-back executed by ONE FULL WARP: 32 predecessors are examined per step (short-lived scan tiles are all in
// the same phase, so a serial walk crosses hundreds of them).  Publishes the aggregate, returns the exclusive prefix in
// every lane and publishes the inclusive one.  Op must be commutative (sum, max).
template <typename Op>
__device__ __forceinline__ u64 lookback_exclusive_warp(u64* chan, size_t tile, u64 aggregate, u32 epoch, Op op) {
    const unsigned lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_relaxed(chan, lb_pack(aggregate, epoch, LB_INCLUSIVE));
        return Op::identity();
    }
    if (lane == 0) st_relaxed(chan + tile, lb_pack(aggregate, epoch, LB_AGGREGATE));
    u64 excl = Op::identity();
    long long t = (long long)tile - 1;  // entry examined by lane 0
    while (true) {
        const long long mine = t - (long long)lane;
        u64 st = LB_INCLUSIVE, payload = Op::identity();  // before tile 0: an inclusive identity
        if (mine >= 0) {
            u64 w;
            do {
                w = ld_relaxed(chan + mine);
                st = lb_state(w, epoch);
            } while (st == LB_NONE);
            payload = lb_payload(w);
        }
        const unsigned incl = __ballot_sync(0xffffffffu, st == LB_INCLUSIVE);
        const unsigned first = incl ? (unsigned)(__ffs(incl) - 1) : 31u;
        u64 v = (lane <= first) ? payload : Op::identity();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
        excl = op(v, excl);
        if (incl) break;
        t -= 32;
    }
    if (lane == 0) st_relaxed(chan + tile, lb_pack(op(excl, aggregate), epoch, LB_INCLUSIVE));
    return excl;
}

}  // namespace psacb200
