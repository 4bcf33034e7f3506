// psac-b200: generalized suffix array of a string set.
//
// Reference: suffix_array::construct_ss (include/suffix_array.hpp:269-363) on a simple_dstringset
// (include/stringset.hpp:33-152): the strings are the maximal runs of non-separator characters of a flat text; the
// result indexes the concatenation WITHOUT separators; a suffix ends with its string (kmer_gen_stringset,
// kmer.hpp:269-355; shift_buckets_ds, shifting.hpp:374-418), identical suffixes of different strings are ordered by
// position (stable tuple sort, :297) and each becomes a bucket of its own (rebucket_gsa*, bucketing.hpp:130-143).
//
// Here: the flat text is sorted WITH its separators as one text in which the separator is code 0 and never matches
// anything -- not even another separator.  That takes three local rules in the kernels of the ordinary construction:
//   * the first sort key of a suffix is cut behind its first separator (gsa_keys_kernel); the sort is stable and fed in
//     text order, so equal cut keys stay in position order;
//   * equal keys that contain a separator are finished buckets of one suffix each, and their LCP is the length up to
//     the separator (resolve_kernel, ResolveArgs::gsa);
//   * later LCP values compare the packed text only up to the first separator (stream_lcp_gsa).
// The separator suffixes themselves sort to the front (their key is 0) in position order, so the rank of "the suffix
// at my string's separator" orders ties between strings exactly like the reference's 0 fill + stable sort, and the
// doubling rounds need no change.  Afterwards the m0 separator suffixes are dropped and text positions are renumbered
// without separators (rank over a separator bit vector).
#pragma once
#include "common.cuh"

namespace psacb200 {

// first sort key of every suffix of the flat text (kbits bits = whole characters), cut behind the first separator
template <typename IdxT>
__global__ void __launch_bounds__(256) gsa_keys_kernel(const u64* __restrict__ stream, u64 n, int lbits, int kbits, u64* __restrict__ keys,
                                                       IdxT* __restrict__ vals) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        keys[i] = gsa_mask_key(stream_extract(stream, i, lbits, kbits), lbits, kbits);
        vals[i] = (IdxT)i;
    }
}

// separator bit vector: bit (i & 63) of bits[i >> 6] <=> text[i] == sep; cnt[w] = separators in word w.  One warp per word.
__global__ void __launch_bounds__(256) gsa_sepbits_kernel(const u8* __restrict__ text, u64 n, u8 sep, u64* __restrict__ bits, u64* __restrict__ cnt,
                                                          u64 nwords) {
    const unsigned lane = threadIdx.x & 31;
    const u64 stride = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwords; w += stride) {
        const u64 i = w * 64 + lane;
        const unsigned lo = __ballot_sync(0xffffffffu, i < n && text[i] == sep);
        const unsigned hi = __ballot_sync(0xffffffffu, i + 32 < n && text[i + 32] == sep);
        if (lane == 0) {
            bits[w] = (u64)lo | ((u64)hi << 32);
            cnt[w] = (u64)(__popc(lo) + __popc(hi));
        }
    }
}

constexpr int GS_THREADS = 256;
constexpr int GS_ITEMS = 8;
constexpr int GS_TILE = GS_THREADS * GS_ITEMS;

// exclusive scan of cnt[] in three steps: tile sums -> tile_scan_kernel (sa_kernels.cuh) -> in-place scan of every tile
__global__ void __launch_bounds__(GS_THREADS) gsa_scan_reduce_kernel(const u64* __restrict__ cnt, u64 nwords, u64* __restrict__ tile_sum) {
    __shared__ u64 s_w[GS_THREADS / 32];
    const u64 base = (u64)blockIdx.x * GS_TILE + (u64)threadIdx.x * GS_ITEMS;
    u64 v = 0;
#pragma unroll
    for (int j = 0; j < GS_ITEMS; ++j) v += (base + j < nwords) ? cnt[base + j] : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < GS_THREADS / 32; ++w) t += s_w[w];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(GS_THREADS) gsa_scan_apply_kernel(u64* __restrict__ cnt, u64 nwords, const u64* __restrict__ tile_pre) {
    __shared__ u64 s_w[GS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 base = (u64)blockIdx.x * GS_TILE + (u64)threadIdx.x * GS_ITEMS;
    u64 c[GS_ITEMS];
    u64 run = 0;
#pragma unroll
    for (int j = 0; j < GS_ITEMS; ++j) {
        c[j] = (base + j < nwords) ? cnt[base + j] : 0;
        run += c[j];
    }
    const u64 incl = warp_inclusive_scan(run, OpSum());
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u64 pre = tile_pre[blockIdx.x] + incl - run;
    for (int w = 0; w < warp; ++w) pre += s_w[w];
#pragma unroll
    for (int j = 0; j < GS_ITEMS; ++j) {
        if (base + j < nwords) cnt[base + j] = pre;
        pre += c[j];
    }
}

// separators before text position i
__device__ __forceinline__ u64 gsa_rank(const u64* __restrict__ bits, const u64* __restrict__ pre, u64 i) {
    const u64 w = i >> 6;
    return pre[w] + (u64)__popcll(bits[w] & ((1ull << (i & 63)) - 1ull));
}

// SA / LCP of the string set: positions [m0, n) of the flat text's arrays, suffixes renumbered without separators
template <typename IdxT, typename OutT>
__global__ void __launch_bounds__(256) gsa_emit_sa_kernel(const IdxT* __restrict__ sa, const IdxT* __restrict__ lcp, u64 m0, u64 n,
                                                          const u64* __restrict__ bits, const u64* __restrict__ pre, OutT* __restrict__ sa_out,
                                                          OutT* __restrict__ lcp_out) {
    for (u64 q = m0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x) {
        const u64 s = (u64)sa[q];
        sa_out[q - m0] = (OutT)(s - gsa_rank(bits, pre, s));
        if (lcp_out != nullptr) lcp_out[q - m0] = (OutT)lcp[q];
    }
}

// ISA of the string set: entries of the non-separator positions, ranks without the m0 separator suffixes in front
template <typename IdxT, typename OutT>
__global__ void __launch_bounds__(256) gsa_emit_isa_kernel(const IdxT* __restrict__ isa, u64 m0, u64 n, const u64* __restrict__ bits,
                                                           const u64* __restrict__ pre, OutT* __restrict__ isa_out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 w = bits[i >> 6];
        if (!((w >> (i & 63)) & 1ull)) isa_out[i - gsa_rank(bits, pre, i)] = (OutT)((u64)isa[i] - m0);
    }
}

}  // namespace psacb200
