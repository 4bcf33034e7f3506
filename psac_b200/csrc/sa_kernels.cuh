// psac-b200: device kernels of the prefix-doubling loop other than the radix sort.
//
// Reference stages restated here as kernels:
//   a2  alphabet histogram            include/alphabet.hpp:48-59          -> byte_hist_kernel
//   a4  k-mer generation              include/kmer.hpp:119-224            -> pack_text_kernel + TextSrc (radix_sort.cuh: fused into digit pass 1)
//   a5  rank shift B2[i] = B[i+h]     include/shifting.hpp:32-122         -> fused into round_keys_kernel (gathers ISA[SA+h])
//   a7  initial k-mer LCP             include/suffix_array.hpp:1353-1396  -> resolve_kernel<FIRST=true>
//   a8  LCP of later rounds           include/suffix_array.hpp:1444-1508  -> resolve_kernel<FIRST=false> (direct compare on packed text)
//   a9  rebucket (head flags + scan)  include/bucketing.hpp:57-123        -> resolve_kernel (warp-shuffle max-scan + look-back)
//   a10 SA->ISA bulk permute          include/bulk_permute.hpp:14-73      -> one radix partition pass by window + isa_scatter_kernel
//                                                                             (direct scatter in resolve_kernel for small / later rounds)
// The engine is not a port: text is packed densely (no sentinel code; suffixes that run past the end are
// ordered by a stable-sort trick, see TextSrc in radix_sort.cuh), the first sort key is ONE word read from the packed
// text instead of a materialised (B1,B2) pair, bucket ids are 0-based head positions (so the final ids ARE the
// ISA and no fix-up pass exists) and later rounds touch only the suffixes that are still in a shared bucket.
#pragma once
#include "common.cuh"

namespace psacb200 {

// ------------------------------------------------------------------ a2: byte histogram
__global__ void __launch_bounds__(512) byte_hist_kernel(const u8* __restrict__ text, size_t n, u64* __restrict__ hist) {
    __shared__ u32 sh[8][256];  // one private histogram per pair of warps cuts same-address contention
    for (int e = threadIdx.x; e < 8 * 256; e += blockDim.x) (&sh[0][0])[e] = 0;
    __syncthreads();
    u32* my = sh[(threadIdx.x >> 6) & 7];
    const size_t head = (16 - ((size_t)text & 15)) & 15;  // bytes before the first 16-byte boundary
    const size_t pre = head < n ? head : n;
    const size_t nvec = (n - pre) / 16;
    const uint4* tv = reinterpret_cast<const uint4*>(text + pre);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 q = __ldcs(tv + i);
        u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) atomicAdd(&my[(w[a] >> (8 * b)) & 0xffu], 1u);
        }
    }
    if (blockIdx.x == 0) {
        for (size_t i = threadIdx.x; i < pre; i += blockDim.x) atomicAdd(&my[text[i]], 1u);
        for (size_t i = pre + nvec * 16 + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&my[text[i]], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        u64 s = 0;
        for (int r = 0; r < 8; ++r) s += sh[r][c];
        if (s) atomicAdd((unsigned long long*)&hist[c], (unsigned long long)s);
    }
}

// ------------------------------------------------------------------ packed text (layout: common.cuh)
__global__ void __launch_bounds__(256) pack_text_kernel(const u8* __restrict__ text, size_t n, CodeTable tab, int lbits, u64* __restrict__ stream,
                                                        size_t nwords) {
    __shared__ u8 lut[256];
    lut[threadIdx.x] = tab.code[threadIdx.x];
    __syncthreads();
    const int cpw = 64 / lbits;
    const bool aligned = (((size_t)text) & 15) == 0;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x) {
        const size_t c0 = w * (size_t)cpw;
        u64 acc = 0;
        if (c0 + cpw <= n && aligned && cpw >= 16) {
            const uint4* tv = reinterpret_cast<const uint4*>(text + c0);
            for (int v = 0; v < cpw / 16; ++v) {
                uint4 q = __ldcs(tv + v);
                u32 ww[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc = (acc << lbits) | lut[(ww[a] >> (8 * b)) & 0xffu];
                }
            }
        } else {
            for (int c = 0; c < cpw; ++c) {
                const size_t i = c0 + c;
                acc = (acc << lbits) | (u64)(i < n ? lut[text[i]] : 0);
            }
        }
        stream[w] = acc;
    }
}

// common characters of two suffixes starting at a+off / b+off in the packed text (used only for the few
// boundaries that are split in rounds >= 1; all the others get their LCP from the sort keys)
// `padded`: the reference's sigma = 256 quirk (include/alphabet.hpp:136,160: the 8-bit code table wraps 0xFF to code
// 0, the code of "past the end"), where its LCP counts matches between 0xFF and the zero padding; then the
// comparison runs over the zero-padded sequences instead of stopping at the end of the shorter suffix.
__device__ __forceinline__ u64 stream_lcp(const u64* __restrict__ stream, u64 n, int lbits, u64 a, u64 b, u64 off, bool padded) {
    const int lsh = 31 - __clz(lbits);  // lbits is 1, 2, 4 or 8: divide by shifting
    const int cpw = 64 >> lsh;
    u64 l = off;
    while (true) {
        const bool ina = a + l < n, inb = b + l < n;
        if (!ina && !inb) break;
        if (!padded && !(ina && inb)) break;
        const u64 wa = ina ? stream_extract(stream, a + l, lbits, 64) : 0;
        const u64 wb = inb ? stream_extract(stream, b + l, lbits, 64) : 0;
        const u64 x = wa ^ wb;
        if (x) {
            l += (u64)(__clzll((long long)x) >> lsh);
            break;
        }
        l += cpw;
    }
    if (padded) return l;
    const u64 la = n - a, lb = n - b;
    const u64 cap = la < lb ? la : lb;
    return l < cap ? l : cap;
}

// string sets (generalized SA): code 0 is the separator and never matches, so the comparison stops at the first separator of
// either suffix as well as at the first difference (the stream is zero behind the text)
__device__ __forceinline__ u64 stream_lcp_gsa(const u64* __restrict__ stream, u64 n, int lbits, u64 a, u64 b, u64 off) {
    const int lsh = 31 - __clz(lbits);
    u64 l = off;
    while (true) {
        const u64 wa = a + l < n ? stream_extract(stream, a + l, lbits, 64) : 0;
        const u64 wb = b + l < n ? stream_extract(stream, b + l, lbits, 64) : 0;
        const u64 stop = (wa ^ wb) | zero_fields(wa, lbits);
        if (stop) {
            l += (u64)(__clzll((long long)stop) >> lsh);
            break;
        }
        l += (u64)(64 >> lsh);
    }
    return l;
}

// ------------------------------------------------------------------ block-distributed ISA in peer-visible memory
// Sharded construction with distributed later rounds (sharded.cuh): block r of the ISA (text positions blk.start(r)...)
// lives in rank r's peer arena as u64; kernels read ("bulk get", reference bulk_rma.hpp:112-135) and write single
// entries straight through the NVLink mappings.  p == 0: not used.
struct PeerIsa {
    u64* blk[16];
    BlkDiv div;
    int p;
    __device__ __forceinline__ u64* at(u64 g) const {
        u64 local;
        const u32 r = div.owner(g, &local);
        return blk[r] + local;
    }
};

// ------------------------------------------------------------------ a7/a8/a9/a10: resolve one round
// Works on the m suffixes that were just sorted (round 0: all n of them, position q; later rounds: the
// unresolved ones, compacted, at SA positions pos[q]).
//   head[q]   = first element of a bucket (sort key differs from the predecessor)
//   bucket[q] = SA position of its head            (inclusive max-scan, warp shuffles + look-back)
//   LCP at every new bucket boundary
// Round 0 (FIRST): the sorted keys are the CARRIED keys of the first sort (the low `drop` bits were consumed by the
//   first digit pass and travelled on as one auxiliary byte per suffix, radix_sort.cuh).  bucket[q] is written sequentially to bucket_out (the SA -> ISA permutation is done afterwards by
//   a partitioned scatter, see isa_scatter_kernel) and, for small inputs only (isa != null), also scattered directly:
//   ISA[suffix[q]] = bucket[q].  The unresolved elements are listed up to a capacity; if more are found (repetitive
//   text) compact_first_kernel lists them from the bucket ids once a large enough buffer exists.
// Later rounds: ISA[suffix[q]] = bucket[q], SA[pos[q]] = suffix[q] directly (few elements), and
//   unresolved' = elements whose bucket still has >= 2 members; their positions and head flags are
//   compacted in order (exclusive sum-scan, look-back) for the next round.
struct ResolveArgs {
    const void* keys;     // sorted keys (KeyC)
    const void* vals;     // sorted suffix indices (IdxT)
    const void* pos_in;   // SA position of element q (rounds >= 1)
    u64 m;                // elements this round
    u64 n;                // text length
    void* sa;             // rounds >= 1: SA[pos[q]] = vals[q]
    void* isa;            // null in round 0 when the partitioned permute follows
    void* bucket_out;     // round 0: bucket id per sorted position (IdxT)
    void* lcp;            // may be null
    void* pos_out;        // compacted positions of the still unresolved elements (first `cap` of them)
    u8* head_out;         // their head flags
    u64 cap;              // capacity of pos_out / head_out (round 0: may be smaller than the number found)
    u64* counts;          // [0] unresolved elements, [1] unresolved buckets (atomic)
    u64* lb_max;          // per tile: aggregate after the reduce phase, exclusive prefix after tile_scan_kernel
    u64* lb_sum;
    const u64* stream;    // packed text
    int lbits;
    int C;                // round 0: characters in the key
    int kbits;            // rounds >= 1: bits of the low key field (rank of suffix+h); the rest is the bucket
    u64 h;                // rounds >= 1: characters already known equal inside a bucket
    int padded_lcp;       // reference quirk: a used character has code 0 and matches the padding (see stream_lcp)
    // ---- sharded construction (sharded.cuh); single GPU: pos_base = 0, no halo, windows = [0, n), suf_out = null
    u64 pos_base;         // round 0: global SA position of local element 0
    const u64* halo;      // round 0: {key, suffix} of the last element of the previous shard, or null
    u64 sa_lo, sa_hi;     // rounds >= 1: SA / LCP positions owned by this shard (sa, lcp point at position sa_lo)
    u64 isa_lo, isa_hi;   // text positions whose ISA entries this shard owns (isa points at entry isa_lo)
    void* suf_out;        // rounds >= 1: suffixes of the still unresolved elements (replicated rounds), or null
    // ---- distributed later rounds: every rank resolves ITS unresolved elements (positions relative to its first SA
    //      position) and puts the new bucket ids into the owners' ISA blocks through peer memory
    PeerIsa pisa;         // pisa.p > 0: ISA[s] = isa_add + bucket goes to pisa.at(s) instead of isa[]
    u64 isa_add;          // first SA position of this rank (bucket ids are global positions)
    WordIdx widx;         // widx.lb > 0: sa[] holds sort words, the suffix is stored as its (owner, local index) field
    void* lcp_main;       // as in HeadsArgs: LCP of positions [main_lo, main_hi) goes to lcp_main[pos], the rest to lcp[pos]
    u64 main_lo, main_hi;
    int lcp_wide;         // rounds >= 1: lcp[] holds 64-bit entries although IdxT is 32 bits wide (see HeadsArgs)
    int gsa;              // string set (reference construct_ss): code 0 separates strings; round-0 keys are cut at their first
                          // separator, equal keys that contain one are finished (rebucket_gsa_kmers, bucketing.hpp:137-143)
};

constexpr int RES_THREADS = 256;
constexpr int RES_ITEMS = 8;
constexpr int RES_TILE = RES_THREADS * RES_ITEMS;

// neighbours of a thread's run: the adjacent lanes hold them (all threads of a warp own consecutive runs); only the
// first / last lane of a warp goes to memory.  Must be called by all 32 lanes.
template <typename T, int N>
__device__ __forceinline__ void load_halo(const T* __restrict__ p, u64 q0, u64 m, u64 (&out)[N + 2]) {
    const unsigned lane = lane_id();
    const u64 left = __shfl_up_sync(0xffffffffu, out[N], 1);
    const u64 right = __shfl_down_sync(0xffffffffu, out[1], 1);
    out[0] = (lane > 0) ? left : ((q0 >= 1 && q0 - 1 < m) ? (u64)p[q0 - 1] : 0);
    out[N + 1] = (lane < 31) ? right : ((q0 + N < m) ? (u64)p[q0 + N] : 0);
}

// out[i] = p[q0 - 1 + i] for i = 0 .. N+1 (0 outside [0, m)); the N inner elements come in as 128-bit loads when the
// run is complete and 16-byte aligned
template <typename T, int N>
__device__ __forceinline__ void load_run(const T* __restrict__ p, u64 q0, u64 m, u64 (&out)[N + 2]) {
    constexpr int PER = 16 / sizeof(T);
    if (q0 + N <= m && ((reinterpret_cast<size_t>(p + q0) & 15) == 0) && (N % PER == 0)) {
        const uint4* v = reinterpret_cast<const uint4*>(p + q0);
#pragma unroll
        for (int c = 0; c < N / PER; ++c) {
            const uint4 t = __ldcs(v + c);
            if (sizeof(T) == 4) {
                out[1 + c * 4 + 0] = t.x;
                out[1 + c * 4 + 1] = t.y;
                out[1 + c * 4 + 2] = t.z;
                out[1 + c * 4 + 3] = t.w;
            } else {
                out[1 + c * 2 + 0] = ((u64)t.y << 32) | t.x;
                out[1 + c * 2 + 1] = ((u64)t.w << 32) | t.z;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) out[1 + i] = (q0 + i < m) ? (u64)p[q0 + i] : 0;
    }
    load_halo<T, N>(p, q0, m, out);
}
// PHASE 0 = reduce: only the tile's aggregates (position of its last head, number of unresolved elements) are written;
// PHASE 1 = apply: reads the tile's exclusive prefixes produced by tile_scan_kernel and writes the results.
// (A single-pass chained scan was measured first: with ~500 k short tiles the look-back waits dominated, 14-17 ms
// against ~5 ms for reduce / scan / apply, profiles/r1_notes.md.)
template <typename IdxT, typename KeyC, bool FIRST, int PHASE>
__global__ void __launch_bounds__(RES_THREADS) resolve_kernel(ResolveArgs A) {
    __shared__ u64 s_wmax[RES_THREADS / 32];
    __shared__ u32 s_wsum[RES_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile = blockIdx.x;
    const u64 q0 = tile * RES_TILE + (u64)tid * RES_ITEMS;
    const IdxT* pos_in = reinterpret_cast<const IdxT*>(A.pos_in);
    const u64 m = A.m, n = A.n;
    const int nbits = A.C * A.lbits;

    // elements q0-1 .. q0+ITEMS (one neighbour on each side)
    u64 key[RES_ITEMS + 2];
    u64 suf[RES_ITEMS + 2];
    load_run<KeyC, RES_ITEMS>(reinterpret_cast<const KeyC*>(A.keys), q0, m, key);
    load_run<IdxT, RES_ITEMS>(reinterpret_cast<const IdxT*>(A.vals), q0, m, suf);
    const bool has_halo = FIRST && A.halo != nullptr;
    if (has_halo && q0 == 0) {  // the element before local position 0 lives on the previous shard
        key[0] = A.halo[0];
        suf[0] = A.halo[1];
    }
    u64 pos[RES_ITEMS];
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        const u64 q = q0 + i;
        pos[i] = FIRST ? (q + A.pos_base) : ((q < m) ? (u64)pos_in[q] : 0);
    }
    // head flags for q0 .. q0+ITEMS (the last one only feeds the "unresolved" test); head[m] = true
    bool head[RES_ITEMS + 1];
    const u64 tail_from = (n >= (u64)A.C) ? (n - (u64)A.C) : 0;  // suffix s runs past the end iff s > tail_from (or n < C)
    const bool all_tail = n < (u64)A.C;
#pragma unroll
    for (int i = 0; i < RES_ITEMS + 1; ++i) {
        const u64 q = q0 + i;
        bool hd;
        if (q >= m || (q == 0 && !has_halo)) {
            hd = true;
        } else {
            hd = key[i + 1] != key[i];
            if (FIRST) hd = hd || all_tail || suf[i + 1] > tail_from || suf[i] > tail_from;
            if (FIRST && A.gsa) hd = hd || (key[i + 1] & ((1ull << A.lbits) - 1ull)) == 0;  // the string ended inside the key
        }
        head[i] = hd;
    }
    // thread-local scans
    u64 mx[RES_ITEMS];
    u32 un[RES_ITEMS];
    u64 run_max = 0;
    u32 run_sum = 0, nbuckets = 0;
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        const bool valid = (q0 + i) < m;
        if (valid && head[i]) run_max = pos[i] > run_max ? pos[i] : run_max;
        mx[i] = run_max;
        const u32 u = (valid && !(head[i] && head[i + 1])) ? 1u : 0u;
        un[i] = run_sum;  // exclusive
        run_sum += u;
        nbuckets += (u && head[i]) ? 1u : 0u;
    }
    // warp + CTA scans
    const u64 wincl_max = warp_inclusive_scan(run_max, OpMax());
    const u32 wincl_sum = warp_inclusive_sum_u32(run_sum);
    u32 wb = nbuckets;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wb += __shfl_xor_sync(0xffffffffu, wb, d);
    if (lane == 31) {
        s_wmax[warp] = wincl_max;
        s_wsum[warp] = wincl_sum;
    }
    u64 texcl_max = __shfl_up_sync(0xffffffffu, wincl_max, 1);
    if (lane == 0) texcl_max = 0;
    const u32 texcl_sum = wincl_sum - run_sum;
    __syncthreads();
    u64 cta_max = 0;
    u32 cta_sum = 0;
    u64 wpre_max = 0;
    u32 wpre_sum = 0;
#pragma unroll
    for (int w = 0; w < RES_THREADS / 32; ++w) {
        if (w == warp) {
            wpre_max = cta_max;
            wpre_sum = cta_sum;
        }
        cta_max = s_wmax[w] > cta_max ? s_wmax[w] : cta_max;
        cta_sum += s_wsum[w];
    }
    if (PHASE == 0) {
        if (tid == 0) {
            A.lb_max[tile] = cta_max;
            A.lb_sum[tile] = cta_sum;
        }
        if (lane == 0 && wb) atomicAdd((unsigned long long*)&A.counts[1], (unsigned long long)wb);
        return;
    }
    u64 pre_max = A.lb_max[tile];  // exclusive prefixes over the preceding tiles (tile_scan_kernel)
    pre_max = wpre_max > pre_max ? wpre_max : pre_max;
    pre_max = texcl_max > pre_max ? texcl_max : pre_max;
    const u64 pre_sum = A.lb_sum[tile] + wpre_sum + texcl_sum;

    IdxT* sa = reinterpret_cast<IdxT*>(A.sa);
    IdxT* isa = reinterpret_cast<IdxT*>(A.isa);
    IdxT* lcp = reinterpret_cast<IdxT*>(A.lcp);
    IdxT* pos_out = reinterpret_cast<IdxT*>(A.pos_out);
    IdxT* bucket_out = reinterpret_cast<IdxT*>(A.bucket_out);
    u64 bucket[RES_ITEMS];
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) bucket[i] = mx[i] > pre_max ? mx[i] : pre_max;
    if (FIRST) {
        // bucket ids leave sequentially: 128-bit stores where the run is complete
        if (q0 + RES_ITEMS <= m && sizeof(IdxT) == 4 && ((reinterpret_cast<size_t>(bucket_out + q0) & 15) == 0)) {
            uint4* o = reinterpret_cast<uint4*>(bucket_out + q0);
#pragma unroll
            for (int c = 0; c < RES_ITEMS / 4; ++c)
                __stcs(o + c, make_uint4((u32)bucket[4 * c], (u32)bucket[4 * c + 1], (u32)bucket[4 * c + 2], (u32)bucket[4 * c + 3]));
        } else {
#pragma unroll
            for (int i = 0; i < RES_ITEMS; ++i)
                if (q0 + i < m) bucket_out[q0 + i] = (IdxT)bucket[i];
        }
    }
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        const u64 q = q0 + i;
        if (q >= m) break;
        const u64 s = suf[i + 1];
        if (!FIRST && A.pisa.p > 0)
            *A.pisa.at(s) = A.isa_add + bucket[i];
        else if (isa != nullptr && s >= A.isa_lo && s < A.isa_hi)
            isa[s - A.isa_lo] = (IdxT)bucket[i];
        const bool own_pos = FIRST || (pos[i] >= A.sa_lo && pos[i] < A.sa_hi);
        if (!FIRST && own_pos) sa[pos[i] - A.sa_lo] = (IdxT)(A.widx.lb > 0 ? A.widx.encode(s) : s);
        if (lcp != nullptr && head[i] && own_pos) {
            if (q == 0 && !has_halo) {
                if (FIRST) lcp[0] = 0;
            } else if (FIRST) {
                const u64 sp = suf[i];
                const u64 x = key[i] ^ key[i + 1];
                u64 c = x ? (u64)(__clzll((long long)(x << (64 - nbits))) >> (31 - __clz(A.lbits))) : (u64)A.C;
                if (A.gsa) {
                    // equal keys: the common prefix ends where the string does (initial_kmer_lcp_gsa, suffix_array.hpp:1404-1441)
                    if (!x) c = key[i] ? (u64)A.C - (u64)((__ffsll((long long)key[i]) - 1) >> (31 - __clz(A.lbits))) : 0;
                } else if (A.padded_lcp) {
                    if (!x) c = stream_lcp(A.stream, n, A.lbits, sp, s, (u64)A.C, true);  // equal keys split by the end-of-text rule
                } else {
                    const u64 la = n - sp, lb = n - s;
                    c = c < la ? c : la;
                    c = c < lb ? c : lb;
                }
                lcp[q] = (IdxT)c;
            } else if ((key[i] >> A.kbits) == (key[i + 1] >> A.kbits)) {
                // same old bucket, different rank of suffix+h: the first h characters agree
                const u64 pr = pos[i] - A.sa_lo;
                IdxT* dst = (A.lcp_main != nullptr && pr >= A.main_lo && pr < A.main_hi) ? reinterpret_cast<IdxT*>(A.lcp_main) : lcp;
                const u64 v = A.gsa ? stream_lcp_gsa(A.stream, n, A.lbits, suf[i], s, A.h) : stream_lcp(A.stream, n, A.lbits, suf[i], s, A.h, A.padded_lcp != 0);
                if (A.lcp_wide)
                    reinterpret_cast<u64*>(dst)[pr] = v;
                else
                    dst[pr] = (IdxT)v;
            }
        }
        const bool unresolved = !(head[i] && head[i + 1]);
        if (unresolved) {
            const u64 o = pre_sum + un[i];
            if (o < A.cap) {
                pos_out[o] = (IdxT)pos[i];
                A.head_out[o] = head[i] ? 1 : 0;
                if (A.suf_out != nullptr) reinterpret_cast<IdxT*>(A.suf_out)[o] = (IdxT)s;
            }
        }
    }
}

// Exclusive scan of the per-tile aggregates of a reduce phase, in place: mx[t] <- max of mx[0..t), sm[t] <- sum of
// sm[0..t); the grand total of sm goes to *total.  One CTA; every thread owns a contiguous slice.
__global__ void __launch_bounds__(1024) tile_scan_kernel(u64* __restrict__ mx, u64* __restrict__ sm, u64 ntiles, u64* __restrict__ total) {
    __shared__ u64 w_max[32], w_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 per = (ntiles + 1023) / 1024;
    const u64 lo = (u64)tid * per, hi = (lo + per < ntiles) ? lo + per : ntiles;
    u64 a_max = 0, a_sum = 0;
    for (u64 t = lo; t < hi; ++t) {
        a_max = mx[t] > a_max ? mx[t] : a_max;
        a_sum += sm[t];
    }
    const u64 i_max = warp_inclusive_scan(a_max, OpMax());
    const u64 i_sum = warp_inclusive_scan(a_sum, OpSum());
    if (lane == 31) {
        w_max[warp] = i_max;
        w_sum[warp] = i_sum;
    }
    u64 e_max = __shfl_up_sync(0xffffffffu, i_max, 1);
    if (lane == 0) e_max = 0;
    u64 e_sum = i_sum - a_sum;
    __syncthreads();
    u64 tot = 0;
    for (int w = 0; w < 32; ++w) {
        if (w < warp) {
            e_max = w_max[w] > e_max ? w_max[w] : e_max;
            e_sum += w_sum[w];
        }
        tot += w_sum[w];
    }
    if (tid == 0 && total != nullptr) *total = tot;
    for (u64 t = lo; t < hi; ++t) {
        const u64 m = mx[t], c = sm[t];
        mx[t] = e_max;
        sm[t] = e_sum;
        e_max = m > e_max ? m : e_max;
        e_sum += c;
    }
}

// ------------------------------------------------------------------ round 0, lean path (a7 + a9)
// Same results as resolve_kernel<FIRST> for 32-bit suffix indices without the padded-LCP quirk, at a fraction of the
// instructions and bytes: heads come from the sorted keys alone.  The only place the suffix indices mattered was the
// end-of-text rule (a suffix shorter than the key is a bucket of its own and caps the LCP), and those <= 63 suffixes
// are located up front: the sort is stable and they were fed first, so tail j sits at lower_bound(its key) + (number
// of shorter tails with the same key).  Reduce (PHASE 0) and apply (PHASE 1) around tile_scan_kernel as before.
struct TailList {
    u32 count;
    u32 len[64];  // length of the suffix (< C)
    u64 pos[64];  // position in the sorted order
};

// (sharded construction: only the tails whose key-prefix bin lies in [bin_lo, bin_hi) are in this shard's sorted range;
//  the others get position ~0 and touch nothing)
// (32-bit carried keys hold the key bits below the top digit; the top digit of position q is its segment in seg_dense)
template <typename KeyC>
__global__ void __launch_bounds__(64) tail_positions_kernel(const KeyC* __restrict__ keys, u64 m, const u64* __restrict__ seg_dense, int seg_shift,
                                                            const u64* __restrict__ stream, u64 n, u64 T, int lbits, int kbits, int pbits, u32 bin_lo,
                                                            u32 bin_hi, TailList* __restrict__ out, int word_shift = 0) {
    __shared__ u64 s_key[64];
    const int j = threadIdx.x;
    u64 full = 0;
    if ((u64)j < T) full = stream_extract(stream, n - 1 - (u64)j, lbits, kbits);
    s_key[j] = full;
    __syncthreads();
    if ((u64)j < T) {
        u64 lo = 0, hi = m;  // first position whose complete key is >= full
        while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            u64 k = (u64)keys[mid] >> word_shift;  // (sharded words: the carried key sits above the suffix index)
            if (seg_dense != nullptr) {
                int a = 0, b = 256;  // last segment starting at or before mid
                while (b - a > 1) {
                    const int c = (a + b) >> 1;
                    if (seg_dense[c] <= mid)
                        a = c;
                    else
                        b = c;
                }
                k |= (u64)a << seg_shift;
            }
            if (k < full)
                lo = mid + 1;
            else
                hi = mid;
        }
        u32 before = 0;
        for (int i = 0; i < j; ++i) before += (s_key[i] == full) ? 1u : 0u;
        const u32 bin = pbits > 0 ? (u32)(full >> (kbits - pbits)) : 0u;
        const bool mine = pbits == 0 || (bin >= bin_lo && bin < bin_hi);
        out->pos[j] = mine ? lo + before : ~0ull;
        out->len[j] = (u32)j + 1;
    }
    if (j == 0) out->count = (u32)T;
}

// marks the heads tiles that touch a segment start p (the tile of p and the tile of p - 1)
__global__ void __launch_bounds__(256) seg_flag_kernel(const u64* __restrict__ seg_dense, u64 m, u64 tile_elems, u8* __restrict__ flags) {
    const u64 p = seg_dense[threadIdx.x];
    if (threadIdx.x > 0 && p > 0 && p < m) {
        flags[p / tile_elems] = 1;
        flags[(p - 1) / tile_elems] = 1;
    }
}

struct HeadsArgs {
    const void* keys;      // sorted carried keys (KeyC)
    const u64* seg_dense;  // 32-bit carried keys: dense start of the 256 top-digit segments (the top digit of a position), else null
    int seg_shift;         // ... and the number of carried bits below the top digit
    const u8* seg_flags;   // ... and per heads tile: 1 if a segment starts inside it or right after it
    const void* vals;      // sorted suffix indices (PosT); read for the direct ISA scatter, the unresolved list and the halo LCP
    u64 m;                 // number of suffixes sorted here
    u64 n;                 // text length
    int lbits, C;
    const TailList* tails;
    void* bucket_out;      // PosT
    void* isa;             // direct scatter (small inputs) or null
    void* lcp;             // or null
    void* pos_out;         // unresolved list (first `cap` entries): SA positions, ...
    u8* head_out;          // ... head flags ...
    void* suf_out;         // ... and suffixes (sharded construction), or null
    u64 cap;
    u64* counts;           // [1] unresolved buckets (atomic); [0] is written by tile_scan_kernel
    u64* agg_max;          // per tile: last head position / exclusive prefix after the scan
    u64* agg_sum;          // per tile: unresolved elements / exclusive prefix
    u64 pos_base;          // sharded construction: SA position of local element 0 (bucket ids and positions are global)
    const u64* halo;       // sharded construction: {key, suffix} of the last element of the previous shard, or null
    void* lcp_main;        // sharded: positions [main_lo, main_hi) of this rank's range lie inside its OWN output block; their LCP
    u64 main_lo, main_hi;  //          goes straight to lcp_main[q] (the caller's block, pre-offset), only the rest to lcp[q]
    int kbits;             // bits of the complete key when it is not C whole characters (0: C * lbits)
    int lcp_wide;          // lcp[] is the caller's array of 64-bit entries although PosT is 32 bits wide (no widening pass later)
    int word_shift;        // WORD: the keys are 64-bit words [carried key | suffix index field]: the key is word >> word_shift,
    WordIdx widx;          //       the suffix index widx.decode(word) (vals is not read)
};

constexpr int HD_THREADS = 256;
constexpr int HD_ITEMS = 16;
constexpr int HD_TILE = HD_THREADS * HD_ITEMS;

template <typename KeyC, typename PosT, int PHASE, bool WORD = false>
__global__ void __launch_bounds__(HD_THREADS) heads_kernel(HeadsArgs A) {
    static_assert(!WORD || sizeof(KeyC) == 8, "words are 64 bits");
    __shared__ u32 s_wmax[HD_THREADS / 32];
    __shared__ u32 s_wsum[HD_THREADS / 32];
    __shared__ TailList s_tails;
    __shared__ int s_has_tail;
    __shared__ u64 s_seg[257];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile = blockIdx.x;
    const u64 t0 = tile * HD_TILE;
    const u64 q0 = t0 + (u64)tid * HD_ITEMS;
    const u64 m = A.m;
    const KeyC* keys = reinterpret_cast<const KeyC*>(A.keys);
    if (tid == 0) s_has_tail = 0;
    // 32-bit carried keys: only the (at most 2 x 255) tiles that touch a segment start need the top digits
    const bool has_seg = A.seg_dense != nullptr && A.seg_flags[tile] != 0;
    if (has_seg)
        for (int e = tid; e <= 256; e += HD_THREADS) s_seg[e] = A.seg_dense[e];
    __syncthreads();
    if (tid < 64 && (u32)tid < A.tails->count) {
        const u64 p = A.tails->pos[tid];
        if (p != ~0ull && p + 1 >= t0 && p <= t0 + HD_TILE) s_has_tail = 1;  // benign race: every writer stores 1
    }
    // ---- my 16 complete keys + one neighbour on each side
    u64 key[HD_ITEMS + 2];
    const bool full_run = q0 + HD_ITEMS <= m;
    if (full_run) {
        if (sizeof(KeyC) == 4) {
            const uint4* v = reinterpret_cast<const uint4*>(keys + q0);
#pragma unroll
            for (int c = 0; c < HD_ITEMS / 4; ++c) {
                const uint4 t = __ldcs(v + c);
                key[1 + 4 * c] = t.x;
                key[2 + 4 * c] = t.y;
                key[3 + 4 * c] = t.z;
                key[4 + 4 * c] = t.w;
            }
        } else {
            const uint4* v = reinterpret_cast<const uint4*>(keys + q0);
#pragma unroll
            for (int c = 0; c < HD_ITEMS / 2; ++c) {
                const uint4 t = __ldcs(v + c);
                key[1 + 2 * c] = ((u64)t.y << 32) | t.x;
                key[2 + 2 * c] = ((u64)t.w << 32) | t.z;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < HD_ITEMS; ++i) {
            const u64 q = q0 + i;
            u64 k = 0;
            if (q < m) k = (u64)keys[q];
            key[1 + i] = k;
        }
    }
    {
        const u64 left = __shfl_up_sync(0xffffffffu, key[HD_ITEMS], 1);
        const u64 right = __shfl_down_sync(0xffffffffu, key[1], 1);
        u64 l = left, r = right;
        if (lane == 0) {
            l = 0;
            if (q0 >= 1 && q0 - 1 < m) {
                l = (u64)keys[q0 - 1];
            } else if (q0 == 0 && A.halo != nullptr) {
                l = A.halo[0];  // last key of the previous shard (only the LCP of position 0 uses it: position 0 is a head)
            }
        }
        if (lane == 31) {
            r = 0;
            if (q0 + HD_ITEMS < m) r = (u64)keys[q0 + HD_ITEMS];
        }
        key[0] = l;
        key[HD_ITEMS + 1] = r;
    }
    if (WORD) {
        // drop the suffix index (the halo key of position 0 is a complete key already)
        const bool halo0 = q0 == 0 && A.halo != nullptr;
#pragma unroll
        for (int i = 0; i < HD_ITEMS + 2; ++i)
            if (!(i == 0 && halo0)) key[i] >>= A.word_shift;
    }
    __syncthreads();
    if (has_seg) {
        // the top digit of a position is the segment it lies in: inside one segment the carried keys alone decide heads
        // and LCPs (the top digits cancel), so this runs only in tiles that contain a segment start.  One binary search
        // for the thread's first element, then a compare per element.
        const u64 qa = q0 >= 1 ? q0 - 1 : 0;
        int sg = 0, hi = 256;
        while (hi - sg > 1) {
            const int c = (sg + hi) >> 1;
            if (s_seg[c] <= qa)
                sg = c;
            else
                hi = c;
        }
#pragma unroll
        for (int i = 0; i < HD_ITEMS + 2; ++i) {
            const u64 q = q0 + i - 1;  // (i = 0 with q0 = 0 is the halo / nothing: its key is not from this array)
            if (q0 + i >= 1 && q < m) {
                while (sg < 255 && s_seg[sg + 1] <= q) ++sg;
                key[i] |= (u64)sg << A.seg_shift;
            }
        }
    }
    // ---- head bits of q0 .. q0+16 (bit 16 only feeds the "unresolved" test)
    u32 head = 0;
#pragma unroll
    for (int i = 0; i <= HD_ITEMS; ++i) {
        const u64 q = q0 + i;
        if (q >= m || q == 0 || key[i + 1] != key[i]) head |= 1u << i;
    }
    if (s_has_tail) {
        if (tid < 64) {
            s_tails.pos[tid] = (u32)tid < A.tails->count ? A.tails->pos[tid] : ~0ull;
            s_tails.len[tid] = (u32)tid < A.tails->count ? A.tails->len[tid] : 0u;
        }
        if (tid == 0) s_tails.count = A.tails->count;
        __syncthreads();
        for (u32 t = 0; t < s_tails.count; ++t) {
            const u64 p = s_tails.pos[t];  // a suffix that runs past the end is a bucket of its own: heads at p and p + 1
            if (p == ~0ull) continue;      // (a tail that sorts into another shard)
            if (p >= q0 && p <= q0 + HD_ITEMS) head |= 1u << (u32)(p - q0);
            if (p + 1 >= q0 && p + 1 <= q0 + HD_ITEMS) head |= 1u << (u32)(p + 1 - q0);
        }
    }
    const u32 valid = (q0 >= m) ? 0u : ((m - q0 >= HD_ITEMS) ? 0xffffu : ((1u << (u32)(m - q0)) - 1u));
    const u32 hv = head & valid;                               // heads among my valid elements
    const u32 unres = ~(head & (head >> 1)) & valid;           // element i is resolved iff it and its successor are heads
    const u32 my_last = hv ? (u32)(q0 + (31 - __clz(hv))) : 0u;  // SA position of my last head (positions fit 32 bits here)
    const u32 my_cnt = __popc(unres);
    // ---- block scans: running max of head positions, running sum of unresolved counts
    u32 imax = my_last, isum = my_cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 om = __shfl_up_sync(0xffffffffu, imax, d);
        const u32 os = __shfl_up_sync(0xffffffffu, isum, d);
        if (lane >= d) {
            imax = om > imax ? om : imax;
            isum += os;
        }
    }
    if (lane == 31) {
        s_wmax[warp] = imax;
        s_wsum[warp] = isum;
    }
    __syncthreads();
    u32 cta_max = 0, cta_sum = 0, wpre_max = 0, wpre_sum = 0;
#pragma unroll
    for (int w = 0; w < HD_THREADS / 32; ++w) {
        if (w == warp) {
            wpre_max = cta_max;
            wpre_sum = cta_sum;
        }
        cta_max = s_wmax[w] > cta_max ? s_wmax[w] : cta_max;
        cta_sum += s_wsum[w];
    }
    if (PHASE == 0) {
        if (tid == 0) {
            A.agg_max[tile] = cta_max;
            A.agg_sum[tile] = cta_sum;
        }
        const u32 nb = __popc(unres & head);  // unresolved buckets start at unresolved heads
        u32 wb = nb;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wb += __shfl_xor_sync(0xffffffffu, wb, d);
        if (lane == 0 && wb) atomicAdd((unsigned long long*)&A.counts[1], (unsigned long long)wb);
        return;
    }
    u32 tex_max = __shfl_up_sync(0xffffffffu, imax, 1);
    if (lane == 0) tex_max = 0;
    u32 run = (u32)A.agg_max[tile];
    run = wpre_max > run ? wpre_max : run;
    run = tex_max > run ? tex_max : run;
    u64 o = A.agg_sum[tile] + wpre_sum + (isum - my_cnt);
    if (q0 >= m) return;
    // ---- bucket ids: position of the latest head
    PosT* bucket_out = reinterpret_cast<PosT*>(A.bucket_out);
    const PosT* vals = reinterpret_cast<const PosT*>(A.vals);
    u32 bucket[HD_ITEMS];
#pragma unroll
    for (int i = 0; i < HD_ITEMS; ++i) {
        if ((hv >> i) & 1u) run = (u32)(q0 + i);
        bucket[i] = run;
    }
    if (bucket_out == nullptr) {
        // (the SA -> ISA step scatters positions and fixes the unresolved suffixes up afterwards)
    } else if (full_run) {
        if (sizeof(PosT) == 4) {
            uint4* ob = reinterpret_cast<uint4*>(bucket_out + q0);
#pragma unroll
            for (int c = 0; c < HD_ITEMS / 4; ++c)
                __stcs(ob + c, make_uint4(bucket[4 * c] + (u32)A.pos_base, bucket[4 * c + 1] + (u32)A.pos_base, bucket[4 * c + 2] + (u32)A.pos_base,
                                          bucket[4 * c + 3] + (u32)A.pos_base));
        } else {
            ulonglong2* ob = reinterpret_cast<ulonglong2*>(bucket_out + q0);
#pragma unroll
            for (int c = 0; c < HD_ITEMS / 2; ++c) __stcs(ob + c, make_ulonglong2(A.pos_base + bucket[2 * c], A.pos_base + bucket[2 * c + 1]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < HD_ITEMS; ++i)
            if (q0 + i < m) bucket_out[q0 + i] = (PosT)(A.pos_base + bucket[i]);
    }
    if (A.isa != nullptr) {
        PosT* isa = reinterpret_cast<PosT*>(A.isa);
#pragma unroll
        for (int i = 0; i < HD_ITEMS; ++i)
            if (q0 + i < m) isa[vals[q0 + i]] = (PosT)(A.pos_base + bucket[i]);
    }
    if (A.lcp != nullptr) {
        const int nbits = A.kbits ? A.kbits : A.C * A.lbits;
        const int lsh = 31 - __clz(A.lbits);  // lbits is 1, 2, 4 or 8: divide by shifting
        u32 l[HD_ITEMS];
#pragma unroll
        for (int i = 0; i < HD_ITEMS; ++i) {
            const u64 x = key[i] ^ key[i + 1];
            l[i] = x ? (u32)(__clzll((long long)(x << (64 - nbits))) >> lsh) : (u32)A.C;
        }
        if (s_has_tail) {
            for (u32 t = 0; t < s_tails.count; ++t) {
                const u64 p = s_tails.pos[t];  // the boundaries on both sides of a short suffix are capped by its length
                if (p == ~0ull) continue;
#pragma unroll
                for (int i = 0; i < HD_ITEMS; ++i)
                    if (q0 + i == p || q0 + i == p + 1) l[i] = l[i] < s_tails.len[t] ? l[i] : s_tails.len[t];
            }
        }
        if (q0 == 0) {
            if (A.halo != nullptr) {
                // boundary with the previous shard: cap by the lengths of both suffixes (either may run past the end)
                const u64 s0 = WORD ? A.widx.decode((u64)keys[0]) : (u64)vals[0];
                const u64 la = A.n - A.halo[1], lb = A.n - s0;
                l[0] = l[0] < la ? l[0] : (u32)la;
                l[0] = l[0] < lb ? l[0] : (u32)lb;
            } else {
                l[0] = 0;
            }
        }
        // every position becomes a head in exactly one round and gets its LCP then: entries of non-heads written here
        // are overwritten by the round that splits them, so the whole run can leave as 128-bit stores
        PosT* lcp = reinterpret_cast<PosT*>(A.lcp);
        bool routed = false;
        if (A.lcp_main != nullptr) {
            PosT* lm = reinterpret_cast<PosT*>(A.lcp_main);
            if (q0 >= A.main_lo && q0 + HD_ITEMS <= A.main_hi) {
                lcp = lm;  // the whole run lies in my own output block
            } else if (q0 + HD_ITEMS > A.main_lo && q0 < A.main_hi) {
                routed = true;  // the run straddles the edge of the block
#pragma unroll
                for (int i = 0; i < HD_ITEMS; ++i)
                    if (q0 + i < m) ((q0 + i >= A.main_lo && q0 + i < A.main_hi) ? lm : lcp)[q0 + i] = (PosT)l[i];
            }
        }
        // A thread's run of 16 entries is 64 or 128 contiguous bytes: written thread by thread, one store instruction of a warp
        // touches 32 different lines, half a sector in each, and the L2 request rate (not DRAM) bounds the kernel.  So a warp
        // whose 512 entries are complete and go to one array transposes them through shared memory (rows padded by 16 bytes:
        // conflict-free both ways) and stores 512 contiguous bytes per instruction.
        bool transposed = false;
        {
            constexpr int HD_ROW_MAX = HD_ITEMS * 8 + 16;
            __shared__ __align__(16) unsigned char s_tr[HD_THREADS / 32][32 * HD_ROW_MAX];
            const int es = (A.lcp_wide || sizeof(PosT) == 8) ? 8 : 4;  // bytes per entry of the destination array
            const u64 w0 = q0 - (u64)lane * HD_ITEMS, w1 = w0 + 32 * HD_ITEMS;
            unsigned char* base = reinterpret_cast<unsigned char*>(A.lcp);
            if (A.lcp_main != nullptr) {
                if (w0 >= A.main_lo && w1 <= A.main_hi)
                    base = reinterpret_cast<unsigned char*>(A.lcp_main);
                else if (!(w1 <= A.main_lo || w0 >= A.main_hi))
                    base = nullptr;  // the warp straddles the edge of the block: thread by thread below
            }
            if (w1 <= m && base != nullptr && ((reinterpret_cast<size_t>(base) + w0 * es) & 15) == 0) {  // (warp-uniform)
                const int row = HD_ITEMS * es + 16;
                unsigned char* mine = s_tr[warp] + lane * row;
                if (es == 8) {
#pragma unroll
                    for (int c = 0; c < HD_ITEMS / 2; ++c) reinterpret_cast<ulonglong2*>(mine)[c] = make_ulonglong2((u64)l[2 * c], (u64)l[2 * c + 1]);
                } else {
#pragma unroll
                    for (int c = 0; c < HD_ITEMS / 4; ++c) reinterpret_cast<uint4*>(mine)[c] = make_uint4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]);
                }
                __syncwarp();
                const int cpt = es;  // 16-byte chunks per thread row (HD_ITEMS * es / 16)
                uint4* dst = reinterpret_cast<uint4*>(base + w0 * es);
                for (int g = lane; g < 32 * cpt; g += 32) {
                    const int t = g / cpt, c = g - t * cpt;
                    __stcs(dst + g, *reinterpret_cast<const uint4*>(s_tr[warp] + t * row + c * 16));
                }
                transposed = true;
            }
        }
        const bool vec_ok = (reinterpret_cast<size_t>(lcp + q0) & 15) == 0;
        if (transposed) {
        } else if (A.lcp_wide) {
            u64* lw = reinterpret_cast<u64*>(A.lcp);
            if (full_run && (reinterpret_cast<size_t>(lw + q0) & 15) == 0) {
                ulonglong2* ol = reinterpret_cast<ulonglong2*>(lw + q0);
#pragma unroll
                for (int c = 0; c < HD_ITEMS / 2; ++c) __stcs(ol + c, make_ulonglong2((u64)l[2 * c], (u64)l[2 * c + 1]));
            } else {
#pragma unroll
                for (int i = 0; i < HD_ITEMS; ++i)
                    if (q0 + i < m) lw[q0 + i] = (u64)l[i];
            }
        } else if (routed) {
        } else if (full_run && vec_ok && sizeof(PosT) == 4) {
            uint4* ol = reinterpret_cast<uint4*>(lcp + q0);
#pragma unroll
            for (int c = 0; c < HD_ITEMS / 4; ++c) __stcs(ol + c, make_uint4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]));
        } else if (full_run && vec_ok) {
            ulonglong2* ol = reinterpret_cast<ulonglong2*>(lcp + q0);
#pragma unroll
            for (int c = 0; c < HD_ITEMS / 2; ++c) __stcs(ol + c, make_ulonglong2((u64)l[2 * c], (u64)l[2 * c + 1]));
        } else {
#pragma unroll
            for (int i = 0; i < HD_ITEMS; ++i)
                if (q0 + i < m) lcp[q0 + i] = (PosT)l[i];
        }
    }
    u32 u = unres;
    while (u) {
        const int i = __ffs(u) - 1;
        u &= u - 1;
        if (o < A.cap) {
            reinterpret_cast<PosT*>(A.pos_out)[o] = (PosT)(A.pos_base + q0 + i);
            A.head_out[o] = (head >> i) & 1u;
            if (A.suf_out != nullptr) {
                if (WORD)
                    reinterpret_cast<u64*>(A.suf_out)[o] = A.widx.decode((u64)keys[q0 + i]);
                else
                    reinterpret_cast<PosT*>(A.suf_out)[o] = vals[q0 + i];
            }
        }
        ++o;
    }
}

// ------------------------------------------------------------------ round 0: list the unresolved elements
// From the bucket ids of round 0 (bucket[q] = position of q's bucket head): q is a head iff bucket[q] == q; it is
// unresolved iff its bucket has >= 2 members.  Writes their positions and head flags in order (sum-scan + look-back).
template <typename IdxT>
__global__ void __launch_bounds__(RES_THREADS) compact_first_kernel(const IdxT* __restrict__ bucket, u64 n, IdxT* __restrict__ pos_out,
                                                                    u8* __restrict__ head_out, u64* __restrict__ lb_sum, u32* __restrict__ tile_counter) {
    __shared__ u32 s_wsum[RES_THREADS / 32];
    __shared__ u64 s_excl_sum;
    __shared__ u32 s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const u64 tile = s_tile;
    const u64 q0 = tile * RES_TILE + (u64)tid * RES_ITEMS;
    bool head[RES_ITEMS + 1];
#pragma unroll
    for (int i = 0; i < RES_ITEMS + 1; ++i) {
        const u64 q = q0 + i;
        head[i] = (q >= n) || ((u64)bucket[q] == q);
    }
    u32 un[RES_ITEMS];
    u32 run_sum = 0;
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        un[i] = run_sum;
        run_sum += ((q0 + i) < n && !(head[i] && head[i + 1])) ? 1u : 0u;
    }
    const u32 wincl = warp_inclusive_sum_u32(run_sum);
    if (lane == 31) s_wsum[warp] = wincl;
    __syncthreads();
    u32 cta_sum = 0, wpre = 0;
#pragma unroll
    for (int w = 0; w < RES_THREADS / 32; ++w) {
        if (w == warp) wpre = cta_sum;
        cta_sum += s_wsum[w];
    }
    if (warp == 0) {
        const u64 e = lookback_exclusive_warp(lb_sum, tile, (u64)cta_sum, 1u, OpSum());
        if (lane == 0) s_excl_sum = e;
    }
    __syncthreads();
    const u64 pre = s_excl_sum + wpre + (wincl - run_sum);
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        const u64 q = q0 + i;
        if (q < n && !(head[i] && head[i + 1])) {
            pos_out[pre + un[i]] = (IdxT)q;
            head_out[pre + un[i]] = head[i] ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------ a10: SA -> ISA, second half of the partitioned permute
// (reference bulk_permute_inplace, include/bulk_permute.hpp:14-73, buckets the pairs by owner rank before scattering;
// here the owner is an L2-sized window of the ISA).  Input: the (suffix, bucket id) pairs partitioned by the top
// bits of the suffix index by one radix pass, so that the CTAs running at any moment scatter into one or two windows of
// n/256 entries that stay resident in L2; every 32-byte sector is written out to HBM once, complete.
template <typename IdxT>
__global__ void __launch_bounds__(256) isa_scatter_kernel(const IdxT* __restrict__ suffix, const IdxT* __restrict__ bucket, IdxT* __restrict__ isa,
                                                          u64 n) {
    constexpr int PER = 16;
    const u64 base = (u64)blockIdx.x * (256 * PER);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const u64 q = base + (u64)i * 256 + threadIdx.x;
        if (q < n) isa[ld_stream(suffix + q)] = ld_stream(bucket + q);
    }
}

// ------------------------------------------------------------------ a5 + key assembly for rounds >= 1
// For unresolved element q (SA position pos[q], suffix s = SA[pos[q]]):
//   key = (index of its bucket's head in the compacted array) << kbits  |  (s+h < n ? ISA[s+h] + 1 : 0)
// so one radix sort orders every bucket by the rank of the suffix h characters further on while keeping
// the buckets where they are.
struct RoundKeyArgs {
    const void* pos;
    const u8* head;
    const void* sa;
    const void* isa;
    u64 m, n, h;
    int kbits;
    u64* keys;
    void* vals;
    u64* lb_max;
    u32* tile_counter;
    const void* suf_in;   // sharded rounds: suffix of element q (instead of sa[pos[q]]), or null
    const u64* rank2;     // sharded rounds: ISA[suffix + h] + 1 (0 past the end) gathered across the shards, or null
    PeerIsa pisa;         // distributed rounds (pisa.p > 0): ISA[suffix + h] is read from the owner's block through peer memory
    u64 pos_add;          // FIXUP: first SA position of this rank
};

// FIXUP (before the first later round, when the SA -> ISA step scattered POSITIONS instead of bucket ids): the step wrote
// every suffix's own position; the unresolved ones must carry the position of their bucket's head instead:
// ISA[suffix] = pos_add + pos[head index] (distributed rounds: through the peer-mapped ISA blocks).
template <typename IdxT, bool FIXUP = false>
__global__ void __launch_bounds__(RES_THREADS) round_keys_kernel(RoundKeyArgs A) {
    __shared__ u64 s_wmax[RES_THREADS / 32];
    __shared__ u64 s_excl_max;
    __shared__ u32 s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(A.tile_counter, 1u);
    __syncthreads();
    const u64 tile = s_tile;
    const u64 q0 = tile * RES_TILE + (u64)tid * RES_ITEMS;
    const IdxT* pos = reinterpret_cast<const IdxT*>(A.pos);
    const IdxT* sa = reinterpret_cast<const IdxT*>(A.sa);
    const IdxT* isa = reinterpret_cast<const IdxT*>(A.isa);
    IdxT* vals = reinterpret_cast<IdxT*>(A.vals);
    u64 mx[RES_ITEMS], suf[RES_ITEMS], k2[RES_ITEMS];
    u64 run_max = 0;
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        const u64 q = q0 + i;
        suf[i] = 0;
        k2[i] = 0;
        if (q < A.m) {
            if (A.head[q]) run_max = q;  // q increases, so "max" is simply the latest head
            const u64 s = A.suf_in != nullptr ? (u64)reinterpret_cast<const IdxT*>(A.suf_in)[q] : (u64)sa[pos[q]];
            suf[i] = s;
            if (FIXUP)
                k2[i] = 0;
            else if (A.pisa.p > 0)
                k2[i] = (s + A.h < A.n) ? *A.pisa.at(s + A.h) + 1 : 0;
            else if (A.rank2 != nullptr)
                k2[i] = A.rank2[q];
            else
                k2[i] = (s + A.h < A.n) ? (u64)isa[s + A.h] + 1 : 0;
        }
        mx[i] = run_max;
    }
    const u64 wincl = warp_inclusive_scan(run_max, OpMax());
    if (lane == 31) s_wmax[warp] = wincl;
    u64 texcl = __shfl_up_sync(0xffffffffu, wincl, 1);
    if (lane == 0) texcl = 0;
    __syncthreads();
    u64 cta_max = 0, wpre = 0;
#pragma unroll
    for (int w = 0; w < RES_THREADS / 32; ++w) {
        if (w == warp) wpre = cta_max;
        cta_max = s_wmax[w] > cta_max ? s_wmax[w] : cta_max;
    }
    if (warp == 0) {
        const u64 e = lookback_exclusive_warp(A.lb_max, tile, cta_max, 1u, OpMax());
        if (lane == 0) s_excl_max = e;
    }
    __syncthreads();
    u64 pre = s_excl_max;
    pre = wpre > pre ? wpre : pre;
    pre = texcl > pre ? texcl : pre;
#pragma unroll
    for (int i = 0; i < RES_ITEMS; ++i) {
        const u64 q = q0 + i;
        if (q >= A.m) break;
        const u64 b = mx[i] > pre ? mx[i] : pre;
        if (FIXUP) {
            if (!A.head[q]) {  // (heads already carry their own position)
                if (A.pisa.p > 0)
                    *A.pisa.at(suf[i]) = A.pos_add + (u64)pos[b];
                else
                    const_cast<IdxT*>(isa)[suf[i]] = (IdxT)(A.pos_add + (u64)pos[b]);
            }
        } else {
            A.keys[q] = (b << A.kbits) | k2[i];
            vals[q] = (IdxT)suf[i];
        }
    }
}

// ------------------------------------------------------------------ output conversion (internal IdxT -> reference index_t)
template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(256) convert_kernel(const SrcT* __restrict__ src, DstT* __restrict__ dst, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) st_stream(dst + i, (DstT)ld_stream(src + i));
}

}  // namespace psacb200
