// psac-b200: host orchestration of the suffix-array / LCP construction and the C ABI (include/psacb200.h).
//
// Path restated: suffix_array<char_t,index_t,_LCP>::construct (reference include/suffix_array.hpp:365-486) at p = 1.
// Loop structure here (see sa_kernels.cuh for how each step maps to the reference's):
//   alphabet -> pack text -> first key (C characters per suffix, one word) -> radix sort -> resolve (buckets, ISA,
//   LCP, compaction of unresolved suffixes) -> while unresolved: keys (bucket, ISA[SA+h]) -> sort -> resolve; h *= 2.
// Output SA / ISA / LCP are the unique arrays of the 0-padded suffix order, hence bit-identical to the reference's
// (SURVEY.md section 0, items 1-2).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/psacb200.h"
#include "radix_sort.cuh"
#include "sa_kernels.cuh"
#include "tree_kernels.cuh"
#include "check_kernels.cuh"
#include "gsa_kernels.cuh"
#include "wide_kernels.cuh"

using namespace psacb200;

static_assert(sizeof(psacb200_stats) == 128, "psacb200_stats layout is mirrored by psac_b200/api.py");

static thread_local std::string g_last_error;
void psacb200::set_last_error(const std::string& msg) { g_last_error = msg; }

namespace {

struct oom_failure {
    size_t bytes;
};
struct arg_failure {
    std::string what;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // grow-only; contents are NOT preserved across a growth
    void reserve(size_t bytes, size_t* total) {
        if (bytes <= cap) return;
        if (p) {
            cudaFree(p);
            *total -= cap;
            p = nullptr;
            cap = 0;
        }
        bytes = align_up(bytes, 1 << 20);
        cudaError_t err = cudaMalloc(&p, bytes);
        if (err != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            throw oom_failure{bytes};
        }
        cap = bytes;
        *total += bytes;
    }
    void release(size_t* total) {
        if (p) cudaFree(p);
        if (total) *total -= cap;
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

enum Phase { PH_H2D, PH_ALPHABET, PH_PACK, PH_HIST, PH_SORT, PH_RESOLVE, PH_ISA, PH_PASS1, PH_ROUNDS, PH_OUTPUT, PH_D2H, PH_TOTAL, PH_COUNT };

}  // namespace

struct psacb200_engine {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    size_t device_bytes = 0;
    uint64_t launches = 0;
    DevBuf text, packed, keys[2], vals[2], vals2, segws, isa, lcp, small, lookback, rk[2], rv[2], rp[2], rh[2], scratch, rep[3], tb[6], alist, gsa[5];
    void* nccl_comm = nullptr;  // ncclComm_t of the sharded construction (sharded.cuh), one rank per engine
    void* nccl_comm2 = nullptr; // a split of it for the copy stream (barrier of the SA -> ISA exchange), or null
    static constexpr int COPY_STREAMS = 4;
    cudaStream_t copy_streams[COPY_STREAMS];  // peer copies of the SA -> ISA exchange (copy engines), overlapping the main stream
    cudaEvent_t ev_cs[COPY_STREAMS];
    cudaEvent_t ev_x[2];
    cudaEvent_t ev_xt[3];  // copy-stream timeline of the exchange: first copy, last copy, barrier
    bool xt_used = false;
    void* peer_map = nullptr;   // PeerArena: peer-visible memory of the sharded construction (sharded.cuh)
    int shard_rank = 0, shard_world = 1;
    u64* h_pinned = nullptr;  // 8192 u64 of pinned host memory for small read-backs and plan uploads
    cudaEvent_t ev_begin[PH_COUNT], ev_end[PH_COUNT];
    bool ev_used[PH_COUNT];
    cudaEvent_t ev_scatter[2 * MAX_PASSES];  // brackets of the scatter kernels of the segmented digit passes
    int scatter_passes = 0;
    psacb200_stats stats;
    u64 selftest_mismatches = 0;
    // fine-grained trace of the last call: consecutive marks on the engine's stream (psacb200_trace)
    static constexpr int TRACE_MAX = 64;
    cudaEvent_t tr_ev[TRACE_MAX];
    const char* tr_name[TRACE_MAX];
    int tr_n = 0;
    void mark(const char* name) {
        if (tr_n < TRACE_MAX) {
            cudaEventRecord(tr_ev[tr_n], stream);
            tr_name[tr_n++] = name;
        }
    }
    bool v1_stats = false;  // sharded v1: the slot of "digit pass 1" holds the key-range selection

    // layout of `small`
    u64* ghist() const { return small.as<u64>(); }
    u64* gbase() const { return small.as<u64>() + MAX_PASSES * RADIX; }
    u64* byte_hist() const { return small.as<u64>() + 2 * MAX_PASSES * RADIX; }
    u64* counts() const { return byte_hist() + 256; }
    u32* counters() const { return reinterpret_cast<u32*>(counts() + 8); }
    u64* shard_meta() const { return reinterpret_cast<u64*>(counters() + 64); }  // 512 u64 of small per-rank exchange data
    void* tail_list() const { return shard_meta() + 512; }                        // TailList (sa_kernels.cuh)
    static size_t small_bytes() { return (2 * MAX_PASSES * RADIX + 256 + 8) * sizeof(u64) + 64 * sizeof(u32) + 512 * sizeof(u64) + 1024; }

    RadixWorkspace radix_ws() const {
        RadixWorkspace ws;
        ws.gbase = gbase();
        ws.tiles = lookback.p;
        ws.tiles_bytes = lookback.cap;
        return ws;
    }
    void begin(Phase p) {
        cudaEventRecord(ev_begin[p], stream);
        ev_used[p] = true;
    }
    void end(Phase p) { cudaEventRecord(ev_end[p], stream); }
    float ms(Phase p) {
        if (!ev_used[p]) return 0.f;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev_begin[p], ev_end[p]) != cudaSuccess) {
            cudaGetLastError();
            return 0.f;
        }
        return t;
    }
};

namespace {

int grid_for(const psacb200_engine* e, size_t work_items, int threads, int per_sm) {
    size_t want = div_up(work_items ? work_items : 1, (size_t)threads);
    size_t cap = (size_t)e->sm_count * per_sm;
    return (int)(want < cap ? want : cap);
}

struct Alphabet {
    uint8_t lut[256];   // reference mapping table (alphabet.hpp:157-164)
    CodeTable dense;    // order-preserving dense codes 0..D-1 used by the packed text
    unsigned sigma = 0;
    unsigned ref_bits = 0;
    int lbits = 1;
    bool zero_code_used = false;  // sigma = 256 quirk: some character shares code 0 with the end-of-text padding
    bool gsa = false;             // string set: dense code 0 is the separator, the characters get 1..D (gsa_kernels.cuh)
};

// reference alphabet<char>::init_mapping_table: codes 1..sigma in byte order, stored in an 8-bit table
void alphabet_from_hist(const u64* hist, Alphabet& a) {
    uint16_t mapped = 1;
    a.sigma = 0;
    for (int c = 0; c < 256; ++c) {
        a.lut[c] = 0;
        if (hist[c]) {
            a.lut[c] = (uint8_t)mapped;  // uchar truncation: 256 -> 0 when every byte value occurs
            ++mapped;
            ++a.sigma;
        }
    }
    a.ref_bits = 0;
    while ((1u << a.ref_bits) < a.sigma + 1) ++a.ref_bits;
}

// dense, order-preserving codes of the characters that occur (order = order of their lut codes)
void dense_codes(const u64* hist, Alphabet& a) {
    bool used[256] = {false};
    for (int c = 0; c < 256; ++c)
        if (hist[c]) used[a.lut[c]] = true;
    a.zero_code_used = used[0];
    int rank[256];
    int d = 0;
    for (int v = 0; v < 256; ++v) rank[v] = used[v] ? d++ : 0;
    if (a.gsa) {
        // (hist[separator] is 0 here: the separator keeps code 0 and matches nothing; a 0 of a caller's LUT cannot be told
        //  from "unused", so the characters must have non-zero LUT codes)
        if (used[0]) throw arg_failure{"string set: the alphabet maps a character that occurs to code 0"};
        for (int c = 0; c < 256; ++c) a.dense.code[c] = hist[c] ? (u8)(rank[a.lut[c]] + 1) : 0;
        d += 1;
    } else {
        for (int c = 0; c < 256; ++c) a.dense.code[c] = hist[c] ? (u8)rank[a.lut[c]] : 0;
    }
    a.lbits = d <= 2 ? 1 : d <= 4 ? 2 : d <= 16 ? 4 : 8;
}

// reduce -> scan -> apply (sa_kernels.cuh): per-tile aggregates, their exclusive scan, then the outputs
template <typename IdxT, typename KeyC>
void launch_resolve(psacb200_engine* e, bool first, const ResolveArgs& A) {
    const u64 ntiles = div_up(A.m, (size_t)RES_TILE);
    PSAC_CUDA(cudaMemsetAsync(A.counts, 0, 2 * sizeof(u64), e->stream));
    if (first) {
        resolve_kernel<IdxT, KeyC, true, 0><<<(unsigned)ntiles, RES_THREADS, 0, e->stream>>>(A);
        tile_scan_kernel<<<1, 1024, 0, e->stream>>>(A.lb_max, A.lb_sum, ntiles, A.counts);
        resolve_kernel<IdxT, KeyC, true, 1><<<(unsigned)ntiles, RES_THREADS, 0, e->stream>>>(A);
    } else {
        resolve_kernel<IdxT, u64, false, 0><<<(unsigned)ntiles, RES_THREADS, 0, e->stream>>>(A);
        tile_scan_kernel<<<1, 1024, 0, e->stream>>>(A.lb_max, A.lb_sum, ntiles, A.counts);
        resolve_kernel<IdxT, u64, false, 1><<<(unsigned)ntiles, RES_THREADS, 0, e->stream>>>(A);
    }
    e->launches += 3;
    PSAC_CUDA(cudaGetLastError());
}

void read_counts(psacb200_engine* e, u64* m, u64* nb) {
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned, e->counts(), 2 * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
    PSAC_CUDA(cudaStreamSynchronize(e->stream));
    *m = e->h_pinned[0];
    *nb = e->h_pinned[1];
}

// Moves an internal result array (IdxT) to the caller's buffer (index_bytes wide, host or device).
template <typename SrcT>
void emit(psacb200_engine* e, const SrcT* src, void* dst, u64 n, int index_bytes, bool dst_is_host) {
    if (dst == nullptr) return;
    if ((const void*)src == dst) {
        if ((int)sizeof(SrcT) == index_bytes) return;  // the result was built in the caller's buffer
        // The internal 32-bit array occupies the lower half of the caller's 64-bit buffer (reserve_buffers): widen in
        // place from the top down.  Elements [c/2, c) are read at bytes [2c, 4c) and written to [4c, 8c): no overlap inside
        // a launch, and everything a launch overwrites has been consumed by the launches before it.
        u64* out = reinterpret_cast<u64*>(dst);
        u64 c1 = n;
        while (c1 > 1) {
            const u64 c0 = (c1 + 1) / 2;  // 8 * c0 >= 4 * c1: the launch writes behind everything it reads
            convert_kernel<SrcT, u64><<<grid_for(e, c1 - c0, 256, 16), 256, 0, e->stream>>>(src + c0, out + c0, c1 - c0);
            e->launches += 1;
            c1 = c0;
        }
        if (c1 == 1) {
            convert_kernel<SrcT, u64><<<1, 32, 0, e->stream>>>(src, out, 1);  // element 0: one thread reads, then writes
            e->launches += 1;
        }
        PSAC_CUDA(cudaGetLastError());
        return;
    }
    if ((int)sizeof(SrcT) == index_bytes) {
        PSAC_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(SrcT), dst_is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, e->stream));
        return;
    }
    // widen / narrow on the device, then move
    void* target = dst;
    if (dst_is_host) {
        e->scratch.reserve(n * (size_t)index_bytes, &e->device_bytes);
        target = e->scratch.p;
    }
    const int grid = grid_for(e, n, 256, 16);
    if (index_bytes == 8)
        convert_kernel<SrcT, u64><<<grid, 256, 0, e->stream>>>(src, reinterpret_cast<u64*>(target), n);
    else
        convert_kernel<SrcT, u32><<<grid, 256, 0, e->stream>>>(src, reinterpret_cast<u32*>(target), n);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    if (dst_is_host) PSAC_CUDA(cudaMemcpyAsync(dst, target, n * (size_t)index_bytes, cudaMemcpyDeviceToHost, e->stream));
}

// Number of characters in the first sort key.
unsigned choose_key_chars(u64 n, int lbits, unsigned k) {
    const unsigned maxC = 64u / (unsigned)lbits;
    if (k != 0) {
        // the reference's first sort key is (kmer_k[i], kmer_k[i+k]) = 2k characters (suffix_array.hpp:381-394)
        u64 c = 2ull * k;
        return (unsigned)std::max<u64>(1, std::min<u64>(c, maxC));
    }
    // Enough characters that random text leaves well under 1 % of the suffixes in shared buckets (n / 2^bits of them),
    // in whole 8-bit digit passes; 40 bits is preferred while it does so because the keys carried after the first
    // digit pass then fit 32 bits (radix_sort.cuh).
    unsigned want = bits_for(n) + 6;
    unsigned nbits = want <= 40 ? 40u : std::min(64u, (want + 7u) / 8u * 8u);
    if (bits_for(n) <= 16) nbits = std::min(nbits, 32u);
    return std::max(1u, std::min(maxC, nbits / (unsigned)lbits));
}

size_t lookback_bytes(u64 n) {
    const size_t rows = suffix_sort_rows(n);  // padded tile rows of the top-digit-first suffix sort
    const size_t seg = align_up(rows * RADIX * sizeof(u32), 256) + (div_up(rows, (size_t)SCAN_CHUNK) + 1) * RADIX * sizeof(u64);
    return std::max(std::max(RadixWorkspace::tiles_bytes_for(n), seg), (size_t)(2 * div_up(n, (size_t)RES_TILE) * sizeof(u64) + div_up(n, (size_t)HD_TILE) + 64));
}

// Device buffers of one construction.  When the caller's outputs are device buffers of the engine's internal index
// width they are used in place (final SA = the value buffer the last digit pass writes, ISA and LCP directly).
// Unresolved-list capacity reserved up front: round 0 lists its unresolved suffixes inside the resolve kernel while
// they fit (random text leaves n / 2^10 of them); a repetitive text that overflows it takes the compact_first_kernel path.
u64 unresolved_cap(u64 n) { return n / 16 + 4096; }

template <typename IdxT>
void reserve_buffers(psacb200_engine* e, u64 n, size_t key_bytes, bool want_lcp, bool ext_sa, bool ext_isa, bool ext_lcp) {
    size_t* tot = &e->device_bytes;
    e->small.reserve(psacb200_engine::small_bytes(), tot);
    e->packed.reserve((n / 8 + 4) * sizeof(u64) + 64, tot);  // worst case 8 bits per character
    if (key_bytes == 4) {
        // top-digit-first sort: padded ping-pong buffers, segment tables, and a third suffix buffer for the dense result
        const size_t pe = suffix_sort_padded_elems(n);
        for (int b = 0; b < 2; ++b) {
            e->keys[b].reserve(pe * sizeof(u32), tot);
            e->vals[b].reserve(pe * sizeof(IdxT), tot);
        }
        if (!ext_sa) e->vals2.reserve(n * sizeof(IdxT), tot);
        e->segws.reserve((2 * (RADIX + 1) + RADIX * RADIX) * sizeof(u64) + suffix_sort_rows(n) * sizeof(u32) + 256, tot);
    } else {
        for (int b = 0; b < 2; ++b) e->keys[b].reserve(n * key_bytes, tot);
        e->vals[0].reserve(n * sizeof(IdxT), tot);
        if (!ext_sa) e->vals[1].reserve(n * sizeof(IdxT), tot);
    }
    if (!ext_isa) e->isa.reserve(n * sizeof(IdxT), tot);
    if (want_lcp && !ext_lcp) e->lcp.reserve(n * sizeof(IdxT), tot);
    e->lookback.reserve(lookback_bytes(n), tot);
    const u64 cap = unresolved_cap(n);
    e->rp[1].reserve(cap * sizeof(IdxT), tot);
    e->rh[1].reserve(cap, tot);
}

template <typename IdxT, typename KeyC>
void construct_core(psacb200_engine* e, u64 n, int index_bytes, unsigned flags, const Alphabet& alpha, unsigned C, void* sa_out, void* isa_out,
                    void* lcp_out, bool out_is_host, bool gsa = false) {
    static_assert(sizeof(KeyC) >= sizeof(IdxT), "bucket ids are staged in a key buffer");
    const bool want_lcp = (flags & PSACB200_LCP) != 0;
    cudaStream_t st = e->stream;
    psacb200_stats& S = e->stats;
    S.internal_index_bytes = sizeof(IdxT);
    S.sort_elt_bytes = (uint32_t)(sizeof(KeyC) + sizeof(IdxT));
    const int lbits = alpha.lbits;
    auto aligned16 = [](const void* q) { return (reinterpret_cast<size_t>(q) & 15) == 0; };  // (the kernels use 128-bit stores)
    // Caller's DEVICE buffers double as the engine's result arrays: directly when the widths match, and as storage for the
    // 32-bit internal array (widened in place at the end, see emit) when the caller's index is 64 bits wide -- at
    // BASELINE configs[3] (2^32 characters) that keeps 48 GiB of device memory free.
    const bool inplace = !out_is_host && (index_bytes == (int)sizeof(IdxT) || (index_bytes == 8 && sizeof(IdxT) == 4)) && aligned16(sa_out) &&
                         aligned16(isa_out) && aligned16(lcp_out);
    const bool ext_sa = inplace, ext_isa = inplace && isa_out != nullptr, ext_lcp = inplace && want_lcp;
    reserve_buffers<IdxT>(e, n, sizeof(KeyC), want_lcp, ext_sa, ext_isa, ext_lcp);

    e->mark("text");
    // ---- first sort (a4 + a6): digit pass 1 reads the packed text, the others the carried keys
    const RadixPlan plan = make_radix_plan(0, (int)C * lbits);
    KeyC* kbuf[2] = {e->keys[0].as<KeyC>(), e->keys[1].as<KeyC>()};
    IdxT* vbuf[2] = {e->vals[0].as<IdxT>(), e->vals[1].as<IdxT>()};
    IdxT* SA = nullptr;
    const u64* seg_dense = nullptr;  // 32-bit carried keys: dense start of every top-digit segment
    uint64_t sort_launches = 0;
    RadixPlan plan_used;
    int x;
    e->begin(PH_SORT);
    if constexpr (sizeof(KeyC) == 4) {
        static_assert(sizeof(IdxT) == 4, "32-bit carried keys go with 32-bit suffix indices");
        SegWorkspace sw;
        sw.seg_dense = e->segws.as<u64>();
        sw.seg_pad = sw.seg_dense + (RADIX + 1);
        sw.segbase = sw.seg_pad + (RADIX + 1);
        sw.tile_info = reinterpret_cast<u32*>(sw.segbase + RADIX * RADIX);
        SA = ext_sa ? reinterpret_cast<IdxT*>(sa_out) : e->vals2.as<IdxT>();
        x = radix_sort_suffixes_msd<IdxT>(e->radix_ws(), sw, e->packed.as<u64>(), n, lbits, (int)C, kbuf, vbuf, SA, st, &plan_used, &sort_launches,
                                          e->ev_end[PH_PASS1], e->ev_scatter);
        e->scatter_passes = plan_used.npass - 1;
        seg_dense = sw.seg_dense;
    } else if (gsa) {
        // string set: keys cut behind the first separator, materialised in text order, then the generic stable LSD sort
        const int fin = plan.npass & 1;
        if (ext_sa) vbuf[fin] = reinterpret_cast<IdxT*>(sa_out), vbuf[1 - fin] = e->vals[0].as<IdxT>();
        gsa_keys_kernel<IdxT><<<grid_for(e, n, 256, 16), 256, 0, st>>>(e->packed.as<u64>(), n, lbits, (int)C * lbits, kbuf[0], vbuf[0]);
        PSAC_CUDA(cudaGetLastError());
        cudaEventRecord(e->ev_end[PH_PASS1], st);
        const bool alt = radix_sort_pairs<u64, IdxT>(e->radix_ws(), kbuf[0], kbuf[1], vbuf[0], vbuf[1], n, 0, (int)C * lbits, st, e->sm_count, &plan_used,
                                                    &sort_launches);
        sort_launches += 1;
        x = alt ? 1 : 0;
        SA = vbuf[x];
    } else {
        const int fin = (plan.npass - 1) & 1;  // buffer index the last pass writes
        if (ext_sa) vbuf[fin] = reinterpret_cast<IdxT*>(sa_out), vbuf[1 - fin] = e->vals[0].as<IdxT>();
        x = radix_sort_suffixes<IdxT>(e->radix_ws(), e->packed.as<u64>(), n, lbits, (int)C, kbuf, vbuf, st, &plan_used, &sort_launches, e->ev_end[PH_PASS1]);
        SA = vbuf[x];
    }
    e->mark("sort");
    e->end(PH_SORT);
    e->launches += sort_launches;
    S.sort_passes = plan_used.npass;
    const int y = 1 - x;
    // scratch suffix buffer of n elements that is free from here on (window partition of the ISA step)
    IdxT* vscratch = sizeof(KeyC) == 4 ? vbuf[0] : vbuf[y];
    const int topbits = plan_used.bits[plan_used.npass - 1];
    const int carried_bits = (int)C * lbits - topbits;  // 32-bit carried keys hold these low bits; the segment is the top digit
    IdxT* ISA = ext_isa ? reinterpret_cast<IdxT*>(isa_out) : e->isa.as<IdxT>();
    IdxT* LCP = want_lcp ? (ext_lcp ? reinterpret_cast<IdxT*>(lcp_out) : e->lcp.as<IdxT>()) : nullptr;

    // ---- resolve round 0 (a7, a9): bucket ids, LCP, count of unresolved suffixes
    const bool partitioned = n >= (1ull << 23);  // below that the ISA fits L2 many times over: scatter directly
    IdxT* bucket = reinterpret_cast<IdxT*>(kbuf[y]);
    e->begin(PH_RESOLVE);
    ResolveArgs R{};
    R.keys = kbuf[x];
    R.vals = SA;
    R.pos_in = nullptr;
    R.m = n;
    R.n = n;
    R.sa = SA;
    R.isa = partitioned ? nullptr : ISA;
    R.bucket_out = bucket;
    R.lcp = LCP;
    R.pos_out = e->rp[1].p;
    R.head_out = e->rh[1].as<u8>();
    R.cap = std::min<u64>(unresolved_cap(n), e->rh[1].cap);
    R.counts = e->counts();
    R.lb_max = e->lookback.as<u64>();
    R.lb_sum = R.lb_max + div_up(n, (size_t)RES_TILE);
    R.stream = e->packed.as<u64>();
    R.lbits = lbits;
    R.C = (int)C;
    R.kbits = 0;
    R.h = 0;
    R.padded_lcp = alpha.zero_code_used ? 1 : 0;
    R.pos_base = 0;
    R.halo = nullptr;
    R.sa_lo = 0;
    R.sa_hi = n;
    R.isa_lo = 0;
    R.isa_hi = n;
    R.suf_out = nullptr;
    R.gsa = gsa ? 1 : 0;
    const bool lean = !gsa && (sizeof(KeyC) == 4 || (sizeof(IdxT) == 4 && (!alpha.zero_code_used || !want_lcp)));
    // 64-bit caller, 32-bit engine: the LCP entries are written 64 bits wide where they are produced (no widening pass)
    const bool lcp_wide = lean && ext_lcp && index_bytes == 8 && sizeof(IdxT) == 4 && !getenv("PSACB200_NO_WIDE_LCP");
    // The SA -> ISA step scatters POSITIONS generated on the fly instead of a bucket-id array (a resolved suffix's bucket id
    // is its position); the unresolved ones are fixed up from their list afterwards.  Saves one written and one read array.
    bool implicit_pos = lean && partitioned && !getenv("PSACB200_NO_IMPLICIT_POS");
    R.lcp_wide = lcp_wide ? 1 : 0;
    HeadsArgs H{};
    if (lean) {
        // lean path: heads from the keys alone (sa_kernels.cuh heads_kernel); the suffixes that run past the end of the
        // text are located in the sorted order first
        const u64 T = (n < (u64)C - 1) ? n : (u64)C - 1;
        static_assert(sizeof(TailList) <= 1024, "TailList must fit its slot in the small buffer");
        TailList* tails = reinterpret_cast<TailList*>(e->tail_list());
        tail_positions_kernel<KeyC><<<1, 64, 0, st>>>(kbuf[x], n, seg_dense, carried_bits, e->packed.as<u64>(), n, T, lbits, (int)C * lbits, 0, 0u, 1u, tails);
        H.keys = kbuf[x];
        H.seg_dense = seg_dense;
        H.seg_shift = carried_bits;
        H.seg_flags = nullptr;
        if (seg_dense != nullptr) {
            u8* flags = e->lookback.as<u8>() + 2 * div_up(n, (size_t)HD_TILE) * sizeof(u64);  // behind the tile aggregates
            PSAC_CUDA(cudaMemsetAsync(flags, 0, div_up(n, (size_t)HD_TILE), st));
            seg_flag_kernel<<<1, 256, 0, st>>>(seg_dense, n, (u64)HD_TILE, flags);
            e->launches += 1;
            H.seg_flags = flags;
        }
        H.vals = SA;
        H.m = n;
        H.n = n;
        H.lbits = lbits;
        H.C = (int)C;
        H.tails = tails;
        H.bucket_out = implicit_pos ? nullptr : bucket;
        H.isa = partitioned ? nullptr : ISA;
        H.lcp = LCP;
        H.lcp_wide = lcp_wide ? 1 : 0;
        H.pos_out = R.pos_out;
        H.head_out = R.head_out;
        H.suf_out = nullptr;
        H.cap = R.cap;
        H.counts = R.counts;
        H.pos_base = 0;
        H.halo = nullptr;
        const u64 ntiles = div_up(n, (size_t)HD_TILE);
        H.agg_max = e->lookback.as<u64>();
        H.agg_sum = H.agg_max + ntiles;
        PSAC_CUDA(cudaMemsetAsync(R.counts, 0, 2 * sizeof(u64), st));
        heads_kernel<KeyC, u32, 0><<<(unsigned)ntiles, HD_THREADS, 0, st>>>(H);
        tile_scan_kernel<<<1, 1024, 0, st>>>(H.agg_max, H.agg_sum, ntiles, R.counts);
        heads_kernel<KeyC, u32, 1><<<(unsigned)ntiles, HD_THREADS, 0, st>>>(H);
        e->launches += 4;
        PSAC_CUDA(cudaGetLastError());
    } else {
        launch_resolve<IdxT, KeyC>(e, true, R);
    }
    e->mark("heads");
    e->end(PH_RESOLVE);
    u64 m = 0, nb = 0;
    read_counts(e, &m, &nb);
    S.unresolved_after_first = m;
    S.rounds = 1;

    // ---- unresolved suffixes (positions + head flags) for the later rounds: listed by the resolve kernel unless they
    //      overflowed its capacity (repetitive text)
    if (m > 0) {
        size_t* tot = &e->device_bytes;
        for (int b = 0; b < 2; ++b) {
            e->rk[b].reserve(m * sizeof(u64), tot);
            e->rv[b].reserve(m * sizeof(IdxT), tot);
        }
        e->rp[0].reserve(m * sizeof(IdxT), tot);
        e->rh[0].reserve(m, tot);
        if (m > R.cap) {
            if (implicit_pos) {
                // repetitive text: the list overflowed, so the bucket ids are needed after all -- apply phase once more
                if constexpr (sizeof(IdxT) == 4) {
                    H.bucket_out = bucket;
                    H.lcp = nullptr;
                    heads_kernel<KeyC, u32, 1><<<(unsigned)div_up(n, (size_t)HD_TILE), HD_THREADS, 0, st>>>(H);
                    e->launches += 1;
                    PSAC_CUDA(cudaGetLastError());
                }
                implicit_pos = false;
            }
            e->rp[1].reserve(m * sizeof(IdxT), tot);
            e->rh[1].reserve(m, tot);
            const u64 ntiles = div_up(n, (size_t)RES_TILE);
            PSAC_CUDA(cudaMemsetAsync(e->lookback.p, 0, ntiles * sizeof(u64), st));
            PSAC_CUDA(cudaMemsetAsync(e->counters() + 18, 0, sizeof(u32), st));
            compact_first_kernel<IdxT><<<(unsigned)ntiles, RES_THREADS, 0, st>>>(bucket, n, e->rp[1].as<IdxT>(), e->rh[1].as<u8>(),
                                                                                 e->lookback.as<u64>(), e->counters() + 18);
            e->launches += 1;
            PSAC_CUDA(cudaGetLastError());
        }
    }

    e->mark("counts");
    // ---- SA -> ISA (a10): partition the (suffix, bucket) pairs by ISA window, then scatter window by window
    e->begin(PH_ISA);
    if (partitioned) {
        const int nbits = (int)bits_for(n - 1);
        const int shift = nbits > RADIX_BITS ? nbits - RADIX_BITS : 0;
        RadixWorkspace ws = e->radix_ws();
        IdxT* part_suffix = vscratch;
        IdxT* part_bucket = reinterpret_cast<IdxT*>(kbuf[x]);  // the sorted keys are dead after resolve
        if (implicit_pos) {
            PosSrc<IdxT, IdxT> src{SA, shift, (u32)(RADIX - 1)};
            launch_pass<PosSrc<IdxT, IdxT>, IdxT, false>(ws, src, part_suffix, part_bucket, nullptr, n, st);
        } else {
            ArraySrc<IdxT, IdxT> src{SA, bucket, nullptr, shift, (u32)(RADIX - 1), (IdxT)0};
            launch_pass<ArraySrc<IdxT, IdxT>, IdxT, false>(ws, src, part_suffix, part_bucket, nullptr, n, st);
        }
        e->mark("isa_window");
        isa_scatter_kernel<IdxT><<<(unsigned)div_up(n, (size_t)4096), 256, 0, st>>>(part_suffix, part_bucket, ISA, n);
        e->launches += LAUNCHES_PER_PASS + 1;
        PSAC_CUDA(cudaGetLastError());
        if (implicit_pos && m > 0) {
            // ISA entries of the unresolved suffixes: the position of their bucket's head (round_keys_kernel FIXUP)
            const u64 ntiles = div_up(m, (size_t)RES_TILE);
            RoundKeyArgs K{};
            K.pos = e->rp[1].p;
            K.head = e->rh[1].as<u8>();
            K.sa = SA;
            K.isa = ISA;
            K.m = m;
            K.n = n;
            K.lb_max = e->lookback.as<u64>();
            K.tile_counter = e->counters() + 17;
            PSAC_CUDA(cudaMemsetAsync(K.lb_max, 0, ntiles * sizeof(u64), st));
            PSAC_CUDA(cudaMemsetAsync(K.tile_counter, 0, sizeof(u32), st));
            round_keys_kernel<IdxT, true><<<(unsigned)ntiles, RES_THREADS, 0, st>>>(K);
            e->launches += 1;
            PSAC_CUDA(cudaGetLastError());
        }
    }
    e->mark("isa_scatter");
    e->end(PH_ISA);

    // ---- later rounds on the unresolved suffixes only (a5, a6, a8, a9, a10, a12)
    if (m > 0) {
        e->begin(PH_ROUNDS);
        const void* pos_in = e->rp[1].p;
        const u8* head_in = e->rh[1].as<u8>();
        const int kbits = (int)bits_for(n);
        u64 h = C;
        int t = 0;
        while (m > 0) {
            const int mbits = (int)bits_for(m - 1);
            if (kbits + mbits > 64) throw arg_failure{"text too repetitive for a 64-bit round key (n >= 2^32 with > 2^30 unresolved suffixes)"};
            const u64 ntiles = div_up(m, (size_t)RES_TILE);
            RoundKeyArgs K{};
            K.pos = pos_in;
            K.head = head_in;
            K.sa = SA;
            K.isa = ISA;
            K.m = m;
            K.n = n;
            K.h = h;
            K.kbits = kbits;
            K.keys = e->rk[0].as<u64>();
            K.vals = e->rv[0].p;
            K.lb_max = e->lookback.as<u64>();
            K.tile_counter = e->counters() + 17;
            PSAC_CUDA(cudaMemsetAsync(K.lb_max, 0, ntiles * sizeof(u64), st));
            PSAC_CUDA(cudaMemsetAsync(K.tile_counter, 0, sizeof(u32), st));
            round_keys_kernel<IdxT><<<(unsigned)ntiles, RES_THREADS, 0, st>>>(K);
            e->launches += 1;
            PSAC_CUDA(cudaGetLastError());
            uint64_t sl = 0;
            const bool alt = radix_sort_pairs<u64, IdxT>(e->radix_ws(), e->rk[0].as<u64>(), e->rk[1].as<u64>(), e->rv[0].as<IdxT>(),
                                                        e->rv[1].as<IdxT>(), m, 0, kbits + mbits, st, e->sm_count, nullptr, &sl);
            e->launches += sl;
            ResolveArgs Q = R;
            Q.keys = e->rk[alt ? 1 : 0].p;
            Q.vals = e->rv[alt ? 1 : 0].p;
            Q.pos_in = pos_in;
            Q.m = m;
            Q.isa = ISA;
            Q.bucket_out = nullptr;
            Q.pos_out = e->rp[t].p;
            Q.head_out = e->rh[t].as<u8>();
            Q.cap = m;
            Q.lb_sum = Q.lb_max + ntiles;
            Q.kbits = kbits;
            Q.h = h;
            launch_resolve<IdxT, u64>(e, false, Q);
            read_counts(e, &m, &nb);
            pos_in = e->rp[t].p;
            head_in = e->rh[t].as<u8>();
            t ^= 1;
            h *= 2;
            S.rounds += 1;
            if (h > 4 * n + 64 && m > 0) throw std::string("prefix doubling did not converge");
        }
        e->end(PH_ROUNDS);
    }

    e->mark("rounds");
    // ---- outputs: SA and ISA (= final bucket ids, 0-based) and LCP in the caller's index width
    e->begin(out_is_host ? PH_D2H : PH_OUTPUT);
    emit<IdxT>(e, SA, sa_out, n, index_bytes, out_is_host);
    emit<IdxT>(e, ISA, isa_out, n, index_bytes, out_is_host);
    if (want_lcp && !lcp_wide) emit<IdxT>(e, LCP, lcp_out, n, index_bytes, out_is_host);
    e->mark("output");
    e->end(out_is_host ? PH_D2H : PH_OUTPUT);
}

// alphabet (a2) + packed text; synchronises once to read the 256-bin histogram
void prepare_text(psacb200_engine* e, const u8* d_text, u64 n, const uint8_t* user_lut, Alphabet& alpha, int gsa_sep = -1, u64* n_sep = nullptr) {
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    e->small.reserve(psacb200_engine::small_bytes(), tot);
    e->packed.reserve((n / 8 + 4) * sizeof(u64) + 64, tot);
    e->begin(PH_ALPHABET);
    PSAC_CUDA(cudaMemsetAsync(e->byte_hist(), 0, 256 * sizeof(u64), st));
    byte_hist_kernel<<<grid_for(e, n / 16 + 1, 512, 4), 512, 0, st>>>(d_text, n, e->byte_hist());
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 16, e->byte_hist(), 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    e->end(PH_ALPHABET);
    PSAC_CUDA(cudaStreamSynchronize(st));
    if (gsa_sep >= 0) {  // string set: the separator is no character of the alphabet (alphabet<char>::from_string of the strings)
        alpha.gsa = true;
        if (n_sep) *n_sep = e->h_pinned[16 + gsa_sep];
        e->h_pinned[16 + gsa_sep] = 0;
    }
    alphabet_from_hist(e->h_pinned + 16, alpha);
    if (user_lut) memcpy(alpha.lut, user_lut, 256);
    dense_codes(e->h_pinned + 16, alpha);
    e->stats.sigma = alpha.sigma;
    e->stats.bits_per_char = alpha.ref_bits;
    e->stats.pack_bits = alpha.lbits;
    e->begin(PH_PACK);
    const int cpw = 64 / alpha.lbits;
    const size_t nwords = div_up(n, (size_t)cpw) + 2;
    pack_text_kernel<<<grid_for(e, nwords, 256, 8), 256, 0, st>>>(d_text, n, alpha.dense, alpha.lbits, e->packed.as<u64>(), nwords);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    e->end(PH_PACK);
}

// ---- device-side certificate (check_kernels.cuh).  counters: e->counts()-adjacent 8 u64 inside `small`
unsigned long long* check_counters(psacb200_engine* e) { return reinterpret_cast<unsigned long long*>(e->shard_meta() + 64); }  // 5 words

void check_launch(psacb200_engine* e, const CheckArgs& A, int index_bytes) {
    const int grid = grid_for(e, A.m, 256, 16);
    if (index_bytes == 4)
        check_sa_kernel<u32><<<grid, 256, 0, e->stream>>>(A);
    else
        check_sa_kernel<u64><<<grid, 256, 0, e->stream>>>(A);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
}

void check_reset(psacb200_engine* e) {
    unsigned long long* bad = check_counters(e);
    PSAC_CUDA(cudaMemsetAsync(bad, 0, 4 * sizeof(u64), e->stream));
    PSAC_CUDA(cudaMemsetAsync(bad + 4, 0xff, sizeof(u64), e->stream));
}

void check_device_core(psacb200_engine* e, const u8* d_text, u64 n, int index_bytes, const void* d_sa, const void* d_isa, const void* d_lcp,
                       psacb200_check_report* rep) {
    memset(rep, 0, sizeof(*rep));
    rep->n = n;
    rep->first_bad = ~0ull;
    rep->checked_lcp = d_lcp != nullptr;
    if (n == 0) return;
    Alphabet alpha;
    prepare_text(e, d_text, n, nullptr, alpha);
    cudaEvent_t e0 = e->ev_begin[PH_OUTPUT], e1 = e->ev_end[PH_OUTPUT];
    PSAC_CUDA(cudaEventRecord(e0, e->stream));
    check_reset(e);
    CheckArgs A{};
    A.sa = d_sa;
    A.lcp = d_lcp;
    A.pos0 = 0;
    A.m = n;
    A.n = n;
    A.halo_sa = 0;
    A.stream = e->packed.as<u64>();
    A.lbits = alpha.lbits;
    A.padded = alpha.zero_code_used ? 1 : 0;
    A.p = 1;
    A.isa_blk[0] = d_isa;
    A.div = BlkDiv::make(n, 1);
    A.bad = check_counters(e);
    check_launch(e, A, index_bytes);
    PSAC_CUDA(cudaEventRecord(e1, e->stream));
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 48, A.bad, 5 * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
    PSAC_CUDA(cudaStreamSynchronize(e->stream));
    rep->bad_range = e->h_pinned[48];
    rep->bad_inverse = e->h_pinned[49];
    rep->bad_order = e->h_pinned[50];
    rep->bad_lcp = e->h_pinned[51];
    rep->first_bad = e->h_pinned[52];
    cudaEventElapsedTime(&rep->ms, e0, e1);
}

// inverse of the dense code table on the device (256 bytes at the end of `small`'s byte-histogram area is too tight: own slot)
u8* upload_inverse_codes(psacb200_engine* e, const Alphabet& alpha, const u64* hist) {
    u8* h = reinterpret_cast<u8*>(e->h_pinned + 7168);  // 256 bytes of the pinned staging area
    memset(h, 0, 256);
    for (int c = 255; c >= 0; --c)
        if (hist[c]) h[alpha.dense.code[c]] = (u8)c;
    e->tb[1].reserve(8192 + 256, &e->device_bytes);
    u8* d = e->tb[1].as<u8>() + 8192;  // (behind the slot of the sharded searcher's table, sharded.cuh dist_search_setup)
    PSAC_CUDA(cudaMemcpyAsync(d, h, 256, cudaMemcpyHostToDevice, e->stream));
    return d;
}

template <typename IdxT>
void lc_launch(psacb200_engine* e, const Alphabet& alpha, const u8* d_inv, u64 n, const void* d_sa, const void* d_lcp, u64 pos0, u64 m, u64 halo_sa, u8* d_lc) {
    if (m == 0) return;
    lc_kernel<IdxT><<<grid_for(e, m, 256, 16), 256, 0, e->stream>>>(e->packed.as<u64>(), alpha.lbits, d_inv, n, reinterpret_cast<const IdxT*>(d_sa),
                                                                    reinterpret_cast<const IdxT*>(d_lcp), pos0, m, halo_sa, d_lc);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
}

void fill_phase_stats(psacb200_engine* e) {
    psacb200_stats& S = e->stats;
    S.device_bytes = e->device_bytes;
    S.ms_total = e->ms(PH_TOTAL);
    S.ms_h2d = e->ms(PH_H2D);
    S.ms_alphabet = e->ms(PH_ALPHABET);
    S.ms_pack = e->ms(PH_PACK);
    S.ms_keygen = 0.f;  // the first key is generated inside digit pass 1
    S.ms_isa = e->ms(PH_ISA);
    {
        float t1 = 0.f;  // digit pass 1 (reads the packed text) runs from the start of PH_SORT to ev_end[PH_PASS1]
        if (e->ev_used[PH_SORT] && cudaEventElapsedTime(&t1, e->ev_begin[PH_SORT], e->ev_end[PH_PASS1]) == cudaSuccess) S.ms_sort_pass1 = t1;
        else cudaGetLastError();
    }
    S.ms_scatter_avg = 0.f;
    if (e->scatter_passes > 0) {
        float sum = 0.f;
        int cnt = 0;
        for (int p = 0; p < e->scatter_passes && p < MAX_PASSES; ++p) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, e->ev_scatter[2 * p], e->ev_scatter[2 * p + 1]) == cudaSuccess) {
                sum += t;
                ++cnt;
            } else {
                cudaGetLastError();
            }
        }
        if (cnt) S.ms_scatter_avg = sum / (float)cnt;
    }
    S.ms_hist = e->ms(PH_HIST);
    S.ms_sort = e->ms(PH_SORT);
    S.ms_resolve = e->ms(PH_RESOLVE);
    S.ms_rounds = e->ms(PH_ROUNDS);
    S.ms_output = e->ms(PH_OUTPUT);
    S.ms_d2h = e->ms(PH_D2H);
    if (e->v1_stats && e->nccl_comm && e->shard_world > 1 && S.ms_sort_pass1 > 0.f)
        S.ms_sort_pass_avg = S.sort_passes ? (S.ms_sort - S.ms_sort_pass1) / (float)S.sort_passes : 0.f;  // sharded: pass1 slot = selection
    else
        S.ms_sort_pass_avg = S.sort_passes > 1 ? (S.ms_sort - S.ms_sort_pass1) / (float)(S.sort_passes - 1) : S.ms_sort;
}

int construct_entry(psacb200_engine* e, const u8* text, bool text_is_host, size_t n, int index_bytes, unsigned flags, unsigned k, const uint8_t* lut,
                    void* sa_out, void* isa_out, void* lcp_out, int out_host = -1) {
    const bool out_is_host = out_host < 0 ? text_is_host : out_host != 0;  // (wide characters: the byte text is on the device already)
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    try {
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        if (n > 0 && (!text || !sa_out)) throw arg_failure{"null text / sa_out"};
        if ((flags & PSACB200_LCP) && n > 0 && !lcp_out) throw arg_failure{"PSACB200_LCP set but lcp_out is null"};
        if (index_bytes == 4 && (u64)n >= (1ull << 32)) throw arg_failure{"32-bit index too small for this text (reference asserts the same)"};
        if ((u64)n >= (1ull << 40)) throw arg_failure{"n too large"};
        PSAC_CUDA(cudaSetDevice(e->device));
        memset(&e->stats, 0, sizeof(e->stats));
        memset(e->ev_used, 0, sizeof(e->ev_used));
        e->scatter_passes = 0;
        e->tr_n = 0;
        e->stats.n = n;
        if (n == 0) return PSACB200_OK;
        e->begin(PH_TOTAL);
        e->mark("begin");
        const u8* d_text = text;
        if (text_is_host) {
            e->begin(PH_H2D);
            e->text.reserve(n + 64, &e->device_bytes);
            PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, n, cudaMemcpyHostToDevice, e->stream));
            e->end(PH_H2D);
            d_text = e->text.as<u8>();
        }
        Alphabet alpha;
        prepare_text(e, d_text, n, lut, alpha);
        const unsigned C = choose_key_chars(n, alpha.lbits, k);
        e->stats.key_chars = C;
        const RadixPlan plan0 = make_radix_plan(0, (int)C * alpha.lbits);
        const int carried_bits = (int)C * alpha.lbits - plan0.bits[plan0.npass - 1];  // key bits below the top digit
        if ((u64)n <= (1ull << 32)) {
            // (the |Sigma| = 256 quirk changes only the LCP -- it counts matches against the zero padding -- so it needs the
            //  generic 64-bit-key path only when the LCP array is wanted; SA / ISA come out of the lean path unchanged:
            //  dense code 0 = 0xFF compares equal to the padding exactly as in the reference's zero-padded k-mers)
            if (carried_bits <= 32 && (!alpha.zero_code_used || !(flags & PSACB200_LCP)))
                construct_core<u32, u32>(e, n, index_bytes, flags, alpha, C, sa_out, isa_out, lcp_out, out_is_host);
            else
                construct_core<u32, u64>(e, n, index_bytes, flags, alpha, C, sa_out, isa_out, lcp_out, out_is_host);
        } else {
            construct_core<u64, u64>(e, n, index_bytes, flags, alpha, C, sa_out, isa_out, lcp_out, out_is_host);
        }
        e->end(PH_TOTAL);
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        fill_phase_stats(e);
        return PSACB200_OK;
    } catch (const cuda_failure& f) {
        set_last_error(std::string("CUDA error: ") + cudaGetErrorString(f.err) + " in " + f.what + " at " + f.file + ":" + std::to_string(f.line));
        cudaGetLastError();
        return PSACB200_ERR_CUDA;
    } catch (const oom_failure& f) {
        set_last_error("out of device memory allocating " + std::to_string(f.bytes) + " bytes");
        return PSACB200_ERR_OOM;
    } catch (const arg_failure& f) {
        set_last_error(f.what);
        return PSACB200_ERR_ARG;
    } catch (const std::string& s) {
        set_last_error(s);
        return PSACB200_ERR_INTERNAL;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return PSACB200_ERR_INTERNAL;
    }
}

template <typename F>
int guarded(F&& f) {
    try {
        return f();
    } catch (const cuda_failure& x) {
        set_last_error(std::string("CUDA error: ") + cudaGetErrorString(x.err) + " in " + x.what + " at " + x.file + ":" + std::to_string(x.line));
        cudaGetLastError();
        return PSACB200_ERR_CUDA;
    } catch (const oom_failure& x) {
        set_last_error("out of device memory allocating " + std::to_string(x.bytes) + " bytes");
        return PSACB200_ERR_OOM;
    } catch (const arg_failure& x) {
        set_last_error(x.what);
        return PSACB200_ERR_ARG;
    } catch (const std::string& s) {
        set_last_error(s);
        return PSACB200_ERR_INTERNAL;
    }
}

// ---- texts over wide characters (wide_kernels.cuh; reference suffix_array<int, ...> with int_alphabet)
template <typename T>
int wide_core(psacb200_engine* e, const void* text, bool text_is_host, size_t n, u32 flip, int index_bytes, unsigned flags, unsigned k, void* sa_out,
              void* isa_out, void* lcp_out, int64_t* distinct_out, uint32_t* n_distinct, bool is_signed) {
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    const T* d_wide = reinterpret_cast<const T*>(text);
    if (text_is_host) {
        e->gsa[0].reserve(n * sizeof(T) + 64, tot);
        PSAC_CUDA(cudaMemcpyAsync(e->gsa[0].p, text, n * sizeof(T), cudaMemcpyHostToDevice, st));
        d_wide = e->gsa[0].as<T>();
    }
    e->gsa[3].reserve((WIDE_SLOTS + 2 + 256) * sizeof(u64), tot);
    unsigned long long* gtab = e->gsa[3].as<unsigned long long>();
    unsigned int* meta = reinterpret_cast<unsigned int*>(gtab + WIDE_SLOTS);
    u32* d_table = reinterpret_cast<u32*>(gtab + WIDE_SLOTS + 2);
    PSAC_CUDA(cudaMemsetAsync(gtab, 0, (WIDE_SLOTS + 2) * sizeof(u64), st));
    wide_distinct_kernel<T><<<grid_for(e, n, 256, 8), 256, 0, st>>>(d_wide, n, flip, gtab, meta);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    u64* h = e->h_pinned;
    PSAC_CUDA(cudaMemcpyAsync(h, gtab, (WIDE_SLOTS + 2) * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    const unsigned int* hm = reinterpret_cast<const unsigned int*>(h + WIDE_SLOTS);
    std::vector<u32> keys;
    for (int i = 0; i < WIDE_SLOTS; ++i)
        if (h[i]) keys.push_back((u32)h[i]);
    if (hm[1] != 0 || keys.size() > (size_t)WIDE_MAX_DISTINCT)
        throw arg_failure{"more than 255 distinct characters: wide-character texts are reduced to one byte per character"};
    std::sort(keys.begin(), keys.end());
    if (n_distinct) *n_distinct = (uint32_t)keys.size();
    if (distinct_out)
        for (size_t i = 0; i < keys.size(); ++i) {
            const u32 raw = keys[i] ^ flip;
            distinct_out[i] = is_signed ? (sizeof(T) == 2 ? (int64_t)(int16_t)raw : (int64_t)(int32_t)raw) : (int64_t)raw;
        }
    u32* hk = reinterpret_cast<u32*>(h);
    for (size_t i = 0; i < keys.size(); ++i) hk[i] = keys[i];
    PSAC_CUDA(cudaMemcpyAsync(d_table, hk, keys.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
    e->text.reserve(n + 64, tot);
    wide_map_kernel<T><<<grid_for(e, n, 256, 8), 256, 0, st>>>(d_wide, n, flip, d_table, (int)keys.size(), e->text.as<u8>());
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    PSAC_CUDA(cudaStreamSynchronize(st));  // (h_pinned is reused by the construction)
    uint8_t lut[256];
    for (int c = 0; c < 256; ++c) lut[c] = (c >= 1 && c <= (int)keys.size()) ? (uint8_t)c : 0;
    const uint64_t launches = e->launches;
    const int rc = construct_entry(e, e->text.as<u8>(), false, n, index_bytes, flags, k, lut, sa_out, isa_out, lcp_out, text_is_host ? 1 : 0);
    (void)launches;
    return rc;
}

int wide_entry(psacb200_engine* e, const void* text, bool text_is_host, size_t n, int char_bytes, int char_signed, int index_bytes, unsigned flags,
               unsigned k, void* sa_out, void* isa_out, void* lcp_out, int64_t* distinct_out, uint32_t* n_distinct) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (char_bytes != 2 && char_bytes != 4) throw arg_failure{"char_bytes must be 2 or 4 (1-byte texts: psacb200_construct)"};
        if (n_distinct) *n_distinct = 0;
        if (n == 0) return PSACB200_OK;
        if (!text || !sa_out) throw arg_failure{"null text / sa_out"};
        PSAC_CUDA(cudaSetDevice(e->device));
        const bool sg = char_signed != 0;
        if (char_bytes == 2)
            return wide_core<uint16_t>(e, text, text_is_host, n, sg ? 0x8000u : 0u, index_bytes, flags, k, sa_out, isa_out, lcp_out, distinct_out, n_distinct, sg);
        return wide_core<uint32_t>(e, text, text_is_host, n, sg ? 0x80000000u : 0u, index_bytes, flags, k, sa_out, isa_out, lcp_out, distinct_out, n_distinct, sg);
    });
}

// ---- generalized suffix array of a string set (gsa_kernels.cuh; reference construct_ss, suffix_array.hpp:269-363)
template <typename IdxT, typename OutT>
void gsa_emit(psacb200_engine* e, u64 len, u64 m0, bool want_lcp, const u64* bits, const u64* pre, void* sa_out, void* isa_out, void* lcp_out,
              bool out_is_host) {
    cudaStream_t st = e->stream;
    const u64 n = len - m0;
    const IdxT* SA = e->gsa[0].as<IdxT>();
    const IdxT* ISA = e->gsa[1].as<IdxT>();
    const IdxT* LCP = want_lcp ? e->gsa[2].as<IdxT>() : nullptr;
    OutT* sa_t = reinterpret_cast<OutT*>(sa_out);
    OutT* lcp_t = want_lcp ? reinterpret_cast<OutT*>(lcp_out) : nullptr;
    OutT* isa_t = reinterpret_cast<OutT*>(isa_out);
    if (out_is_host) {
        e->scratch.reserve(2 * n * sizeof(OutT) + 64, &e->device_bytes);
        sa_t = e->scratch.as<OutT>();
        lcp_t = want_lcp ? sa_t + n : nullptr;
        isa_t = sa_t;
    }
    gsa_emit_sa_kernel<IdxT, OutT><<<grid_for(e, n, 256, 16), 256, 0, st>>>(SA, LCP, m0, len, bits, pre, sa_t, lcp_t);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    if (out_is_host) {
        PSAC_CUDA(cudaMemcpyAsync(sa_out, sa_t, n * sizeof(OutT), cudaMemcpyDeviceToHost, st));
        if (want_lcp) PSAC_CUDA(cudaMemcpyAsync(lcp_out, lcp_t, n * sizeof(OutT), cudaMemcpyDeviceToHost, st));
    }
    if (isa_out != nullptr) {
        gsa_emit_isa_kernel<IdxT, OutT><<<grid_for(e, len, 256, 16), 256, 0, st>>>(ISA, m0, len, bits, pre, isa_t);
        e->launches += 1;
        PSAC_CUDA(cudaGetLastError());
        if (out_is_host) PSAC_CUDA(cudaMemcpyAsync(isa_out, isa_t, n * sizeof(OutT), cudaMemcpyDeviceToHost, st));
    }
}

template <typename IdxT>
void gsa_core(psacb200_engine* e, const u8* d_text, u64 len, u64 m0, u8 sep, int index_bytes, unsigned flags, const Alphabet& alpha, unsigned C,
              void* sa_out, void* isa_out, void* lcp_out, bool out_is_host) {
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    const bool want_lcp = (flags & PSACB200_LCP) != 0;
    // SA / ISA / LCP of the flat text (separators included) in engine buffers of the internal index width
    for (int b = 0; b < (want_lcp ? 3 : 2); ++b) e->gsa[b].reserve(len * sizeof(IdxT) + 64, tot);
    construct_core<IdxT, u64>(e, len, (int)sizeof(IdxT), flags, alpha, C, e->gsa[0].p, e->gsa[1].p, want_lcp ? e->gsa[2].p : nullptr, false, true);
    // separator bit vector + separators before every 64-character word
    e->begin(out_is_host ? PH_D2H : PH_OUTPUT);
    const u64 nwords = div_up(len, (size_t)64), ntiles = div_up(nwords, (size_t)GS_TILE);
    e->gsa[3].reserve(nwords * sizeof(u64), tot);
    e->gsa[4].reserve((nwords + 2 * ntiles + 8) * sizeof(u64), tot);
    u64* bits = e->gsa[3].as<u64>();
    u64* pre = e->gsa[4].as<u64>();
    u64* tile_sum = pre + nwords;
    u64* tile_dummy = tile_sum + ntiles;  // tile_scan_kernel scans a max array along with the sums
    gsa_sepbits_kernel<<<grid_for(e, nwords * 32, 256, 16), 256, 0, st>>>(d_text, len, sep, bits, pre, nwords);
    gsa_scan_reduce_kernel<<<(unsigned)ntiles, GS_THREADS, 0, st>>>(pre, nwords, tile_sum);
    PSAC_CUDA(cudaMemsetAsync(tile_dummy, 0, ntiles * sizeof(u64), st));
    tile_scan_kernel<<<1, 1024, 0, st>>>(tile_dummy, tile_sum, ntiles, nullptr);
    gsa_scan_apply_kernel<<<(unsigned)ntiles, GS_THREADS, 0, st>>>(pre, nwords, tile_sum);
    e->launches += 4;
    PSAC_CUDA(cudaGetLastError());
    if (index_bytes == 8)
        gsa_emit<IdxT, u64>(e, len, m0, want_lcp, bits, pre, sa_out, isa_out, lcp_out, out_is_host);
    else
        gsa_emit<IdxT, u32>(e, len, m0, want_lcp, bits, pre, sa_out, isa_out, lcp_out, out_is_host);
    e->mark("gsa_output");
    e->end(out_is_host ? PH_D2H : PH_OUTPUT);
}

int gsa_entry(psacb200_engine* e, const u8* text, bool text_is_host, size_t len, uint8_t sep, int index_bytes, unsigned flags, const uint8_t* lut,
              void* sa_out, void* isa_out, void* lcp_out, uint64_t* n_out) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        if (!n_out) throw arg_failure{"null n_out"};
        if (len > 0 && (!text || !sa_out)) throw arg_failure{"null text / sa_out"};
        if ((flags & PSACB200_LCP) && len > 0 && !lcp_out) throw arg_failure{"PSACB200_LCP set but lcp_out is null"};
        if (index_bytes == 4 && (u64)len + 1 >= (1ull << 32)) throw arg_failure{"32-bit index too small for this text"};
        if ((u64)len >= (1ull << 40)) throw arg_failure{"n too large"};
        PSAC_CUDA(cudaSetDevice(e->device));
        memset(&e->stats, 0, sizeof(e->stats));
        memset(e->ev_used, 0, sizeof(e->ev_used));
        e->scatter_passes = 0;
        e->tr_n = 0;
        *n_out = 0;
        if (len == 0) return PSACB200_OK;
        e->begin(PH_TOTAL);
        e->mark("begin");
        // The flat text must END with a separator: the end of the text would otherwise rank before every separator in the
        // doubling rounds, while the last string's suffixes belong behind their equals (ties are ordered by position).
        const u8* d_text = text;
        bool append = false;
        if (text_is_host) {
            append = text[len - 1] != sep;
            e->begin(PH_H2D);
            e->text.reserve(len + 64, &e->device_bytes);
            PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, len, cudaMemcpyHostToDevice, e->stream));
            e->end(PH_H2D);
            d_text = e->text.as<u8>();
        } else {
            u8* last = reinterpret_cast<u8*>(e->h_pinned);
            PSAC_CUDA(cudaMemcpyAsync(last, text + len - 1, 1, cudaMemcpyDeviceToHost, e->stream));
            PSAC_CUDA(cudaStreamSynchronize(e->stream));
            append = *last != sep;
            if (append) {
                e->text.reserve(len + 64, &e->device_bytes);
                PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, len, cudaMemcpyDeviceToDevice, e->stream));
                d_text = e->text.as<u8>();
            }
        }
        if (append) {
            PSAC_CUDA(cudaMemsetAsync(e->text.as<u8>() + len, (int)sep, 1, e->stream));
            len += 1;
        }
        Alphabet alpha;
        u64 m0 = 0;
        prepare_text(e, d_text, len, lut, alpha, (int)sep, &m0);
        const u64 n = (u64)len - m0;
        *n_out = n;
        e->stats.n = n;
        if (n > 0) {
            // separators use a code of their own, so only part of the code space occurs in a key: take the widest key
            const unsigned C = std::max(1u, ((u64)len >= 65536 ? 64u : 32u) / (unsigned)alpha.lbits);
            e->stats.key_chars = C;
            if ((u64)len <= (1ull << 32))
                gsa_core<u32>(e, d_text, len, m0, sep, index_bytes, flags, alpha, C, sa_out, isa_out, lcp_out, text_is_host);
            else
                gsa_core<u64>(e, d_text, len, m0, sep, index_bytes, flags, alpha, C, sa_out, isa_out, lcp_out, text_is_host);
        }
        e->end(PH_TOTAL);
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        fill_phase_stats(e);
        return PSACB200_OK;
    });
}

template <typename KeyT>
bool sort_dispatch_val(psacb200_engine* e, void* k, void* ka, void* v, void* va, size_t n, int val_bytes, int b0, int b1, uint64_t* sl) {
    RadixWorkspace ws = e->radix_ws();
    if (val_bytes == 0)
        return radix_sort_pairs<KeyT, NoVal>(ws, (KeyT*)k, (KeyT*)ka, (NoVal*)nullptr, (NoVal*)nullptr, n, b0, b1, e->stream, e->sm_count, nullptr, sl);
    if (val_bytes == 4) return radix_sort_pairs<KeyT, u32>(ws, (KeyT*)k, (KeyT*)ka, (u32*)v, (u32*)va, n, b0, b1, e->stream, e->sm_count, nullptr, sl);
    return radix_sort_pairs<KeyT, u64>(ws, (KeyT*)k, (KeyT*)ka, (u64*)v, (u64*)va, n, b0, b1, e->stream, e->sm_count, nullptr, sl);
}

}  // namespace

// ---- ANSV: min-tree levels + search kernel (tree_kernels.cuh)
namespace {
// list of the positions the tile kernels leave to the exact search (random text: 3 % of them; room for all)
AnsvList ansv_list(psacb200_engine* e, u64 m) {
    AnsvList L;
    L.cap = m + 64;
    e->alist.reserve(L.cap * sizeof(u64) + 64, &e->device_bytes);
    L.entries = e->alist.as<u64>();
    L.count = reinterpret_cast<unsigned long long*>(e->shard_meta() + 150);
    return L;
}
template <typename T>
void ansv_device(psacb200_engine* e, const T* d_vals, u64 n, int left_type, int right_type, u64 nonsv, u64* d_left, u64* d_right) {
    MinTree<T> t{};
    t.level[0] = d_vals;
    t.size[0] = n;
    t.levels = 1;
    // upper levels: minima of blocks of 32, until one block is left
    u64 total = 0;
    for (u64 m = n; m > ANSV_FAN;) {
        m = div_up(m, (size_t)ANSV_FAN);
        total += m;
    }
    e->tb[0].reserve((total + 1) * sizeof(T), &e->device_bytes);
    T* base = e->tb[0].as<T>();
    while (t.size[t.levels - 1] > ANSV_FAN) {
        if (t.levels >= ANSV_MAX_LEVELS) throw arg_failure{"ANSV input too large"};
        const u64 m = div_up(t.size[t.levels - 1], (size_t)ANSV_FAN);
        mintree_level_kernel<T><<<grid_for(e, m, 256, 8), 256, 0, e->stream>>>(t.level[t.levels - 1], t.size[t.levels - 1], base, m);
        e->launches += 1;
        t.level[t.levels] = base;
        t.size[t.levels] = m;
        base += m;
        t.levels += 1;
    }
    LocalSearch<T> sr{t};
    launch_ansv_tile<T, LocalSearch<T>>(sr, d_vals, 0, n, left_type, right_type, nonsv, d_left, d_right, ansv_list(e, n), e->sm_count, e->stream);
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
}
}  // namespace

#include "sharded.cuh"

// ================================================================================================ C ABI
extern "C" {

const char* psacb200_last_error(void) { return g_last_error.c_str(); }

int psacb200_create(int device, psacb200_engine** out) {
    if (!out) {
        set_last_error("null out pointer");
        return PSACB200_ERR_ARG;
    }
    *out = nullptr;
    return guarded([&]() -> int {
        int count = 0;
        cudaError_t err = cudaGetDeviceCount(&count);
        if (err != cudaSuccess || count == 0) {
            cudaGetLastError();
            set_last_error(std::string("no CUDA device available (") + cudaGetErrorString(err) + "); psacb200 has no CPU fallback");
            return PSACB200_ERR_CUDA;
        }
        if (device < 0 || device >= count) throw arg_failure{"device ordinal out of range"};
        PSAC_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        PSAC_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw arg_failure{std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; this library is built for sm_100a only"};
        psacb200_engine* e = new psacb200_engine();
        e->device = device;
        e->sm_count = prop.multiProcessorCount;
        PSAC_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        for (int i = 0; i < psacb200_engine::COPY_STREAMS; ++i) {
            PSAC_CUDA(cudaStreamCreateWithFlags(&e->copy_streams[i], cudaStreamNonBlocking));
            PSAC_CUDA(cudaEventCreateWithFlags(&e->ev_cs[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 2; ++i) PSAC_CUDA(cudaEventCreateWithFlags(&e->ev_x[i], cudaEventDisableTiming));
        for (int i = 0; i < 3; ++i) PSAC_CUDA(cudaEventCreate(&e->ev_xt[i]));
        PSAC_CUDA(cudaMallocHost((void**)&e->h_pinned, 8192 * sizeof(u64)));
        for (int i = 0; i < PH_COUNT; ++i) {
            PSAC_CUDA(cudaEventCreate(&e->ev_begin[i]));
            PSAC_CUDA(cudaEventCreate(&e->ev_end[i]));
            e->ev_used[i] = false;
        }
        for (int i = 0; i < 2 * MAX_PASSES; ++i) PSAC_CUDA(cudaEventCreate(&e->ev_scatter[i]));
        for (int i = 0; i < psacb200_engine::TRACE_MAX; ++i) PSAC_CUDA(cudaEventCreate(&e->tr_ev[i]));
        memset(&e->stats, 0, sizeof(e->stats));
        e->small.reserve(psacb200_engine::small_bytes(), &e->device_bytes);
        // hardware self-test of the ranking assumption of the radix passes (radix_sort.cuh): refuse to run if it fails
        unsigned long long* bad = reinterpret_cast<unsigned long long*>(e->counts());
        PSAC_CUDA(cudaMemsetAsync(bad, 0, sizeof(u64), e->stream));
        atoms_order_selftest_kernel<16><<<2 * e->sm_count, 512, 0, e->stream>>>(1u, 256u, bad);
        atoms_order_selftest_kernel<16><<<2 * e->sm_count, 512, 0, e->stream>>>(2u, 5u, bad);
        e->launches += 2;
        PSAC_CUDA(cudaMemcpyAsync(e->h_pinned, bad, sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        e->selftest_mismatches = e->h_pinned[0];
        // ranking mode of the digit passes: the single-ATOMS ranking only where the self-test confirms the lane order;
        // otherwise (or with PSACB200_SAFE_RANK=1) the match.any ranking built on documented primitives only
        const char* force = getenv("PSACB200_SAFE_RANK");
        if (e->h_pinned[0] != 0 || (force && force[0] == '1')) g_safe_rank = true;
        *out = e;
        return PSACB200_OK;
    });
}

void psacb200_destroy(psacb200_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    DevBuf* all[] = {&e->text, &e->packed, &e->keys[0], &e->keys[1], &e->vals[0], &e->vals[1], &e->vals2, &e->segws, &e->isa, &e->lcp, &e->small, &e->lookback,
                     &e->rk[0], &e->rk[1], &e->rv[0], &e->rv[1], &e->rp[0], &e->rp[1], &e->rh[0], &e->rh[1], &e->scratch, &e->rep[0], &e->rep[1], &e->rep[2], &e->tb[0], &e->tb[1], &e->tb[2], &e->tb[3], &e->tb[4], &e->tb[5], &e->alist, &e->gsa[0], &e->gsa[1], &e->gsa[2], &e->gsa[3], &e->gsa[4]};
    for (DevBuf* b : all) b->release(nullptr);
    for (int i = 0; i < PH_COUNT; ++i) {
        cudaEventDestroy(e->ev_begin[i]);
        cudaEventDestroy(e->ev_end[i]);
    }
    for (int i = 0; i < 2 * MAX_PASSES; ++i) cudaEventDestroy(e->ev_scatter[i]);
    for (int i = 0; i < psacb200_engine::TRACE_MAX; ++i) cudaEventDestroy(e->tr_ev[i]);
    if (e->peer_map) {
        // (psacb200_comm_finalize is the collective, ordered release; this is the local last resort)
        arena_release(e, nullptr);
        delete reinterpret_cast<PeerArena*>(e->peer_map);
    }
    if (e->nccl_comm2 && g_nccl.CommDestroy) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(e->nccl_comm2));
    if (e->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(e->nccl_comm));
    for (int i = 0; i < 2; ++i) cudaEventDestroy(e->ev_x[i]);
    for (int i = 0; i < 3; ++i) cudaEventDestroy(e->ev_xt[i]);
    for (int i = 0; i < psacb200_engine::COPY_STREAMS; ++i) {
        cudaStreamDestroy(e->copy_streams[i]);
        cudaEventDestroy(e->ev_cs[i]);
    }
    if (e->h_pinned) cudaFreeHost(e->h_pinned);
    cudaStreamDestroy(e->stream);
    delete e;
}

uint64_t psacb200_launch_count(const psacb200_engine* e) { return e ? e->launches : 0; }

int psacb200_trace(psacb200_engine* e, char* buf, size_t buf_len) {
    if (!e || !buf || buf_len == 0) return PSACB200_ERR_ARG;
    std::string out;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (int i = 1; i < e->tr_n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e->tr_ev[i - 1], e->tr_ev[i]) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        char tmp[96];
        snprintf(tmp, sizeof(tmp), "%s=%.3f;", e->tr_name[i], t);
        out += tmp;
    }
    if (e->xt_used && e->tr_n > 0) {
        cudaStreamSynchronize(e->copy_streams[0]);
        float a = 0.f, b = 0.f, c = 0.f;
        if (cudaEventElapsedTime(&a, e->tr_ev[0], e->ev_xt[0]) == cudaSuccess && cudaEventElapsedTime(&b, e->ev_xt[0], e->ev_xt[1]) == cudaSuccess &&
            cudaEventElapsedTime(&c, e->ev_xt[1], e->ev_xt[2]) == cudaSuccess) {
            char tmp[160];
            snprintf(tmp, sizeof(tmp), "x_start_at=%.3f;x_copies=%.3f;x_barrier=%.3f;", a, b, c);
            out += tmp;
        } else {
            cudaGetLastError();
        }
    }
    snprintf(buf, buf_len, "%s", out.c_str());
    return PSACB200_OK;
}

int psacb200_rank_mode(const psacb200_engine* e, uint64_t* selftest_mismatches) {
    if (e && selftest_mismatches) *selftest_mismatches = e->selftest_mismatches;
    return g_safe_rank ? 1 : 0;
}

void* psacb200_stream(const psacb200_engine* e) { return e ? (void*)e->stream : nullptr; }

int psacb200_get_stats(const psacb200_engine* e, psacb200_stats* out) {
    if (!e || !out) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    *out = e->stats;
    return PSACB200_OK;
}

int psacb200_reserve(psacb200_engine* e, size_t n, int index_bytes, unsigned flags) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        PSAC_CUDA(cudaSetDevice(e->device));
        // sized for the automatic key length (k = 0) and host outputs (the superset of the device-output case)
        const bool lcp = (flags & PSACB200_LCP) != 0;
        if ((u64)n <= (1ull << 32))
            reserve_buffers<u32>(e, n, sizeof(u32), lcp, false, false, false);  // 32-bit carried keys (key <= 40 bits)
        else
            reserve_buffers<u64>(e, n, sizeof(u64), lcp, false, false, false);
        e->text.reserve(n + 64, &e->device_bytes);
        if ((u64)n <= (1ull << 32) && index_bytes == 8) e->scratch.reserve(n * 8, &e->device_bytes);
        return PSACB200_OK;
    });
}

int psacb200_alphabet(psacb200_engine* e, const uint8_t* text, size_t n, uint8_t lut[256], uint32_t* sigma, uint32_t* bits_per_char) {
    if (!e || (!text && n) || !lut) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        PSAC_CUDA(cudaSetDevice(e->device));
        e->text.reserve(n + 64, &e->device_bytes);
        PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, n, cudaMemcpyHostToDevice, e->stream));
        PSAC_CUDA(cudaMemsetAsync(e->byte_hist(), 0, 256 * sizeof(u64), e->stream));
        if (n) {
            byte_hist_kernel<<<grid_for(e, n / 16 + 1, 512, 4), 512, 0, e->stream>>>(e->text.as<u8>(), n, e->byte_hist());
            e->launches += 1;
        }
        PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 16, e->byte_hist(), 256 * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        Alphabet a;
        alphabet_from_hist(e->h_pinned + 16, a);
        memcpy(lut, a.lut, 256);
        if (sigma) *sigma = a.sigma;
        if (bits_per_char) *bits_per_char = a.ref_bits;
        return PSACB200_OK;
    });
}

int psacb200_construct(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, unsigned flags, unsigned k, void* sa_out, void* isa_out,
                       void* lcp_out) {
    return construct_entry(e, text, true, n, index_bytes, flags, k, nullptr, sa_out, isa_out, lcp_out);
}

int psacb200_construct_alphabet(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, unsigned flags, unsigned k, const uint8_t lut[256],
                                void* sa_out, void* isa_out, void* lcp_out) {
    return construct_entry(e, text, true, n, index_bytes, flags, k, lut, sa_out, isa_out, lcp_out);
}

int psacb200_construct_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, unsigned flags, unsigned k, void* d_sa, void* d_isa,
                              void* d_lcp) {
    return construct_entry(e, d_text, false, n, index_bytes, flags, k, nullptr, d_sa, d_isa, d_lcp);
}

int psacb200_construct_wide(psacb200_engine* e, const void* text, size_t n, int char_bytes, int char_signed, int index_bytes, unsigned flags, unsigned k,
                            void* sa_out, void* isa_out, void* lcp_out, int64_t* distinct_out, uint32_t* n_distinct) {
    return wide_entry(e, text, true, n, char_bytes, char_signed, index_bytes, flags, k, sa_out, isa_out, lcp_out, distinct_out, n_distinct);
}

int psacb200_construct_wide_device(psacb200_engine* e, const void* d_text, size_t n, int char_bytes, int char_signed, int index_bytes, unsigned flags,
                                   unsigned k, void* d_sa, void* d_isa, void* d_lcp, int64_t* distinct_out, uint32_t* n_distinct) {
    return wide_entry(e, d_text, false, n, char_bytes, char_signed, index_bytes, flags, k, d_sa, d_isa, d_lcp, distinct_out, n_distinct);
}

int psacb200_construct_ss(psacb200_engine* e, const uint8_t* flat, size_t len, uint8_t sep, int index_bytes, unsigned flags, const uint8_t* lut,
                          void* sa_out, void* isa_out, void* lcp_out, uint64_t* n_out) {
    return gsa_entry(e, flat, true, len, sep, index_bytes, flags, lut, sa_out, isa_out, lcp_out, n_out);
}

int psacb200_construct_ss_device(psacb200_engine* e, const uint8_t* d_flat, size_t len, uint8_t sep, int index_bytes, unsigned flags, const uint8_t* lut,
                                 void* d_sa, void* d_isa, void* d_lcp, uint64_t* n_out) {
    return gsa_entry(e, d_flat, false, len, sep, index_bytes, flags, lut, d_sa, d_isa, d_lcp, n_out);
}

int psacb200_sort_pairs(psacb200_engine* e, void* d_keys, void* d_keys_alt, void* d_vals, void* d_vals_alt, size_t n, int key_bytes, int val_bytes,
                        int begin_bit, int end_bit) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if ((key_bytes != 4 && key_bytes != 8) || (val_bytes != 0 && val_bytes != 4 && val_bytes != 8)) throw arg_failure{"unsupported key/value width"};
        if (begin_bit < 0 || end_bit > key_bytes * 8 || begin_bit > end_bit) throw arg_failure{"bad bit range"};
        PSAC_CUDA(cudaSetDevice(e->device));
        e->lookback.reserve(RadixWorkspace::tiles_bytes_for(n), &e->device_bytes);
        uint64_t sl = 0;
        bool alt = key_bytes == 8 ? sort_dispatch_val<u64>(e, d_keys, d_keys_alt, d_vals, d_vals_alt, n, val_bytes, begin_bit, end_bit, &sl)
                                  : sort_dispatch_val<u32>(e, d_keys, d_keys_alt, d_vals, d_vals_alt, n, val_bytes, begin_bit, end_bit, &sl);
        e->launches += sl;
        if (alt) {
            PSAC_CUDA(cudaMemcpyAsync(d_keys, d_keys_alt, n * (size_t)key_bytes, cudaMemcpyDeviceToDevice, e->stream));
            if (val_bytes) PSAC_CUDA(cudaMemcpyAsync(d_vals, d_vals_alt, n * (size_t)val_bytes, cudaMemcpyDeviceToDevice, e->stream));
        }
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}

int psacb200_sort_pairs_host(psacb200_engine* e, void* keys, void* vals, size_t n, int key_bytes, int val_bytes, int begin_bit, int end_bit) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    int rc = guarded([&]() -> int {
        PSAC_CUDA(cudaSetDevice(e->device));
        size_t* tot = &e->device_bytes;
        e->rk[0].reserve(n * 8, tot);
        e->rk[1].reserve(n * 8, tot);
        e->rv[0].reserve(n * 8, tot);
        e->rv[1].reserve(n * 8, tot);
        PSAC_CUDA(cudaMemcpyAsync(e->rk[0].p, keys, n * (size_t)key_bytes, cudaMemcpyHostToDevice, e->stream));
        if (val_bytes) PSAC_CUDA(cudaMemcpyAsync(e->rv[0].p, vals, n * (size_t)val_bytes, cudaMemcpyHostToDevice, e->stream));
        return PSACB200_OK;
    });
    if (rc) return rc;
    rc = psacb200_sort_pairs(e, e->rk[0].p, e->rk[1].p, e->rv[0].p, e->rv[1].p, n, key_bytes, val_bytes, begin_bit, end_bit);
    if (rc) return rc;
    return guarded([&]() -> int {
        PSAC_CUDA(cudaMemcpyAsync(keys, e->rk[0].p, n * (size_t)key_bytes, cudaMemcpyDeviceToHost, e->stream));
        if (val_bytes) PSAC_CUDA(cudaMemcpyAsync(vals, e->rv[0].p, n * (size_t)val_bytes, cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}


// ---- sharded construction over the GPUs of one box (sharded.cuh)
int psacb200_comm_unique_id(uint8_t id[128]) {
    return guarded([&]() -> int {
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        g_nccl.load();
        ncclUniqueId u;
        PSAC_NCCL(g_nccl.GetUniqueId(&u));
        memcpy(id, &u, 128);
        return PSACB200_OK;
    });
}

int psacb200_comm_init(psacb200_engine* e, const uint8_t id[128], int rank, int world) {
    if (!e || !id) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (world < 1 || world > 16 || rank < 0 || rank >= world) throw arg_failure{"bad rank / world size (1..16 ranks of one box)"};
        PSAC_CUDA(cudaSetDevice(e->device));
        g_nccl.load();
        if (e->nccl_comm2) {
            g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(e->nccl_comm2));
            e->nccl_comm2 = nullptr;
        }
        if (e->nccl_comm) {
            g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(e->nccl_comm));
            e->nccl_comm = nullptr;
        }
        ncclUniqueId u;
        memcpy(&u, id, 128);
        ncclComm_t c;
        PSAC_NCCL(g_nccl.CommInitRank(&c, world, u, rank));
        e->nccl_comm = c;
        if (g_nccl.CommSplit && world > 1) {
            ncclComm_t c2 = nullptr;
            if (g_nccl.CommSplit(c, 0, rank, &c2, nullptr) == ncclSuccess) e->nccl_comm2 = c2;
        }
        if (!e->peer_map && !getenv("PSACB200_NO_PEER")) e->peer_map = new PeerArena();  // PSACB200_NO_PEER=1: NCCL all-to-all-v instead of peer stores
        e->shard_rank = rank;
        e->shard_world = world;
        return PSACB200_OK;
    });
}

int psacb200_construct_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, unsigned flags, unsigned k,
                               void* d_sa_local, void* d_isa_local, void* d_lcp_local) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (!e->nccl_comm) throw arg_failure{"psacb200_comm_init has not been called on this engine"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        if (index_bytes == 4 && (u64)n_global >= (1ull << 32)) throw arg_failure{"32-bit index too small for this text (reference asserts the same)"};
        if (n_local > 0 && (!d_text_local || !d_sa_local)) throw arg_failure{"null text / sa_out"};
        if ((flags & PSACB200_LCP) && n_local > 0 && !d_lcp_local) throw arg_failure{"PSACB200_LCP set but lcp_out is null"};
        PSAC_CUDA(cudaSetDevice(e->device));
        ShardComm C{reinterpret_cast<ncclComm_t>(e->nccl_comm), e->shard_rank, e->shard_world};
        const u64 n = n_global;
        if (n == 0) return PSACB200_OK;
        bool sharded = (C.world > 1 || getenv("PSACB200_FORCE_SHARDED")) && n >= (u64)C.world * (1ull << 16);  // (forced at one rank: profiling)
        if (sharded) {
            memset(&e->stats, 0, sizeof(e->stats));
            memset(e->ev_used, 0, sizeof(e->ev_used));
            e->scatter_passes = 0;
            e->tr_n = 0;
            e->xt_used = false;
            e->stats.n = n;
            e->begin(PH_TOTAL);
            e->mark("begin");
            e->v1_stats = false;
            sharded = construct_sharded_v2(e, C, d_text_local, n_local, n, index_bytes, flags, k, d_sa_local, d_isa_local, d_lcp_local);
            e->stats.sharded_scheme = sharded ? 2u : 0u;
            if (!sharded) {
                sharded = construct_sharded_core(e, C, d_text_local, n_local, n, index_bytes, flags, k, d_sa_local, d_isa_local, d_lcp_local);
                e->stats.sharded_scheme = sharded ? 1u : 0u;
            }
            e->end(PH_TOTAL);
            PSAC_CUDA(cudaStreamSynchronize(e->stream));
            if (sharded) {
                fill_phase_stats(e);
                return PSACB200_OK;
            }
            if (n > (1ull << 32)) throw arg_failure{"sharded construction: text too skewed / repetitive for this round's sharded scheme and too large to build replicated"};
        }
        {
            // too small (or too skewed) to shard: every GPU builds the whole index from the all-gathered text and keeps its block
            gather_text(e, C, d_text_local, n_local, n);
            const BlkDist blk(n, C.world);
            const bool lcp = (flags & PSACB200_LCP) != 0;
            size_t* tot = &e->device_bytes;
            for (int i = 0; i < 3; ++i)
                if (i < 2 || lcp) e->rep[i].reserve(n * (size_t)index_bytes + 64, tot);
            int rc = construct_entry(e, e->text.as<u8>(), false, n, index_bytes, flags, k, nullptr, e->rep[0].p, e->rep[1].p, lcp ? e->rep[2].p : nullptr);
            if (rc != PSACB200_OK) return rc;
            const size_t o = blk.start(C.rank) * (size_t)index_bytes, bytes = n_local * (size_t)index_bytes;
            if (bytes) {
                PSAC_CUDA(cudaMemcpyAsync(d_sa_local, e->rep[0].as<u8>() + o, bytes, cudaMemcpyDeviceToDevice, e->stream));
                if (d_isa_local) PSAC_CUDA(cudaMemcpyAsync(d_isa_local, e->rep[1].as<u8>() + o, bytes, cudaMemcpyDeviceToDevice, e->stream));
                if (lcp) PSAC_CUDA(cudaMemcpyAsync(d_lcp_local, e->rep[2].as<u8>() + o, bytes, cudaMemcpyDeviceToDevice, e->stream));
            }
            PSAC_CUDA(cudaStreamSynchronize(e->stream));
            return PSACB200_OK;
        }
    });
}

int psacb200_comm_finalize(psacb200_engine* e) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        PSAC_CUDA(cudaSetDevice(e->device));
        if (e->nccl_comm) {
            ShardComm C{reinterpret_cast<ncclComm_t>(e->nccl_comm), e->shard_rank, e->shard_world};
            arena_release(e, &C);
        }
        return PSACB200_OK;
    });
}

// ---- device-side certificate (check_kernels.cuh)
int psacb200_check_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, const void* d_sa, const void* d_isa, const void* d_lcp,
                          psacb200_check_report* report) {
    if (!e || !report) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        if (n > 0 && (!d_text || !d_sa || !d_isa)) throw arg_failure{"null text / sa / isa"};
        PSAC_CUDA(cudaSetDevice(e->device));
        check_device_core(e, d_text, n, index_bytes, d_sa, d_isa, d_lcp, report);
        return PSACB200_OK;
    });
}

int psacb200_check(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, const void* sa, const void* isa, const void* lcp,
                   psacb200_check_report* report) {
    if (!e || !report) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        if (n > 0 && (!text || !sa || !isa)) throw arg_failure{"null text / sa / isa"};
        PSAC_CUDA(cudaSetDevice(e->device));
        size_t* tot = &e->device_bytes;
        const size_t bytes = n * (size_t)index_bytes;
        e->text.reserve(n + 64, tot);
        for (int i = 0; i < 3; ++i)
            if (i < 2 || lcp) e->rep[i].reserve(bytes + 64, tot);
        if (n) {
            PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, n, cudaMemcpyHostToDevice, e->stream));
            PSAC_CUDA(cudaMemcpyAsync(e->rep[0].p, sa, bytes, cudaMemcpyHostToDevice, e->stream));
            PSAC_CUDA(cudaMemcpyAsync(e->rep[1].p, isa, bytes, cudaMemcpyHostToDevice, e->stream));
            if (lcp) PSAC_CUDA(cudaMemcpyAsync(e->rep[2].p, lcp, bytes, cudaMemcpyHostToDevice, e->stream));
        }
        check_device_core(e, e->text.as<u8>(), n, index_bytes, e->rep[0].p, e->rep[1].p, lcp ? e->rep[2].p : nullptr, report);
        return PSACB200_OK;
    });
}

int psacb200_check_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, const void* d_sa_local,
                           const void* d_isa_local, const void* d_lcp_local, psacb200_check_report* report) {
    if (!e || !report) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (!e->nccl_comm) throw arg_failure{"psacb200_comm_init has not been called on this engine"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        if (n_local > 0 && (!d_text_local || !d_sa_local || !d_isa_local)) throw arg_failure{"null text / sa / isa"};
        PSAC_CUDA(cudaSetDevice(e->device));
        ShardComm C{reinterpret_cast<ncclComm_t>(e->nccl_comm), e->shard_rank, e->shard_world};
        if (C.world == 1) {
            check_device_core(e, d_text_local, n_global, index_bytes, d_sa_local, d_isa_local, d_lcp_local, report);
            return PSACB200_OK;
        }
        check_sharded_core(e, C, d_text_local, n_local, n_global, index_bytes, d_sa_local, d_isa_local, d_lcp_local, report);
        return PSACB200_OK;
    });
}

// ---- left-branching characters Lc (check_kernels.cuh lc_kernel)
int psacb200_lc_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, const void* d_sa, const void* d_lcp, uint8_t* d_lc) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    return guarded([&]() -> int {
        if (!d_text || !d_sa || !d_lcp || !d_lc) throw arg_failure{"null argument"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        PSAC_CUDA(cudaSetDevice(e->device));
        Alphabet alpha;
        prepare_text(e, d_text, n, nullptr, alpha);
        const u8* d_inv = upload_inverse_codes(e, alpha, e->h_pinned + 16);
        if (index_bytes == 4)
            lc_launch<u32>(e, alpha, d_inv, n, d_sa, d_lcp, 0, n, 0, d_lc);
        else
            lc_launch<u64>(e, alpha, d_inv, n, d_sa, d_lcp, 0, n, 0, d_lc);
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}

int psacb200_lc(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, const void* sa, const void* lcp, uint8_t* lc) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    return guarded([&]() -> int {
        if (!text || !sa || !lcp || !lc) throw arg_failure{"null argument"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        PSAC_CUDA(cudaSetDevice(e->device));
        size_t* tot = &e->device_bytes;
        const size_t bytes = n * (size_t)index_bytes;
        e->text.reserve(n + 64, tot);
        e->rep[0].reserve(bytes + 64, tot);
        e->rep[1].reserve(bytes + 64, tot);
        e->rep[2].reserve(n + 64, tot);
        PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, n, cudaMemcpyHostToDevice, e->stream));
        PSAC_CUDA(cudaMemcpyAsync(e->rep[0].p, sa, bytes, cudaMemcpyHostToDevice, e->stream));
        PSAC_CUDA(cudaMemcpyAsync(e->rep[1].p, lcp, bytes, cudaMemcpyHostToDevice, e->stream));
        const int rc = psacb200_lc_device(e, e->text.as<u8>(), n, index_bytes, e->rep[0].p, e->rep[1].p, e->rep[2].as<u8>());
        if (rc != PSACB200_OK) return rc;
        PSAC_CUDA(cudaMemcpyAsync(lc, e->rep[2].p, n, cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}

int psacb200_lc_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, const void* d_sa_local,
                        const void* d_lcp_local, uint8_t* d_lc_local) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (!e->nccl_comm) throw arg_failure{"psacb200_comm_init has not been called on this engine"};
        if (n_local && (!d_text_local || !d_sa_local || !d_lcp_local || !d_lc_local)) throw arg_failure{"null argument"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        PSAC_CUDA(cudaSetDevice(e->device));
        if (n_global == 0) return PSACB200_OK;
        ShardComm C{reinterpret_cast<ncclComm_t>(e->nccl_comm), e->shard_rank, e->shard_world};
        const BlkDist blk(n_global, C.world);
        if (blk.size(C.rank) != n_local) throw arg_failure{"the arrays must be equally block decomposed across all ranks (reference suffix_array.hpp:226)"};
        Alphabet alpha;
        prepare_text_sharded(e, C, d_text_local, n_local, n_global, alpha);
        const u8* d_inv = upload_inverse_codes(e, alpha, e->h_pinned + 16);
        // SA[pos0 - 1]: the last SA element of the previous non-empty rank
        u64* d_last = e->shard_meta();
        PSAC_CUDA(cudaMemsetAsync(d_last + C.rank, 0, sizeof(u64), e->stream));
        if (n_local)
            PSAC_CUDA(cudaMemcpyAsync(d_last + C.rank, reinterpret_cast<const u8*>(d_sa_local) + (n_local - 1) * (size_t)index_bytes, (size_t)index_bytes,
                                      cudaMemcpyDeviceToDevice, e->stream));
        PSAC_NCCL(g_nccl.AllGather(d_last + C.rank, d_last, 1, ncclUint64, C.comm, e->stream));
        PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 64, d_last, (size_t)C.world * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        u64 halo = 0;
        for (int r = C.rank - 1; r >= 0; --r)
            if (blk.size(r)) {
                halo = e->h_pinned[64 + r];
                break;
            }
        if (index_bytes == 4)
            lc_launch<u32>(e, alpha, d_inv, n_global, d_sa_local, d_lcp_local, blk.start(C.rank), n_local, halo, d_lc_local);
        else
            lc_launch<u64>(e, alpha, d_inv, n_global, d_sa_local, d_lcp_local, blk.start(C.rank), n_local, halo, d_lc_local);
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}

// host-side plans of the sharded construction, exposed for the CPU tests (no GPU needed)
void psacb200_blk_dist(uint64_t n, int p, int r, uint64_t* start, uint64_t* size) {
    BlkDist b(n, p);
    if (start) *start = b.start(r);
    if (size) *size = b.size(r);
}

int psacb200_choose_splitters(const uint64_t* hist, size_t nbins, uint64_t n, int p, uint64_t* first_out, uint64_t* count_out) {
    if (!hist || !first_out || !count_out || p < 1) return PSACB200_ERR_ARG;
    std::vector<size_t> first;
    std::vector<u64> count;
    choose_splitters(hist, nbins, n, p, first, count);
    for (int r = 0; r <= p; ++r) first_out[r] = first[r];
    for (int r = 0; r < p; ++r) count_out[r] = count[r];
    return PSACB200_OK;
}


int psacb200_plan_word_exchange(const uint64_t* cnt, int p, int nb, uint64_t n, uint64_t pad_tile, uint64_t* first, uint64_t* cnt_key, int32_t* owner,
                                uint64_t* seg_dense, uint64_t* seg_pad, uint64_t* run_off, int* balanced) {
    if (!cnt || p < 1 || p > 16 || nb < 1 || nb > 256 || !first || !cnt_key || !owner || !seg_dense || !seg_pad || !run_off || !balanced) return PSACB200_ERR_ARG;
    WordExchangePlan P;
    plan_word_exchange(cnt, p, nb, n, pad_tile ? pad_tile : 1, P);
    for (int r = 0; r <= p; ++r) first[r] = P.first[r];
    for (int r = 0; r < p; ++r) cnt_key[r] = P.cnt_key[r];
    for (int d = 0; d < nb; ++d) owner[d] = P.owner[d];
    memcpy(seg_dense, P.seg_dense.data(), P.seg_dense.size() * sizeof(u64));
    memcpy(seg_pad, P.seg_pad.data(), P.seg_pad.size() * sizeof(u64));
    memcpy(run_off, P.run_off.data(), P.run_off.size() * sizeof(u64));
    *balanced = P.balanced ? 1 : 0;
    return PSACB200_OK;
}

// ---- ANSV and suffix tree (tree_kernels.cuh)
int psacb200_ansv(psacb200_engine* e, const void* vals, size_t n, int val_bytes, int left_type, int right_type, uint64_t nonsv, uint64_t* left,
                  uint64_t* right) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    return guarded([&]() -> int {
        if (!vals || !left || !right) throw arg_failure{"null argument"};
        if (val_bytes != 4 && val_bytes != 8) throw arg_failure{"val_bytes must be 4 or 8"};
        if (left_type < 0 || left_type > 2 || right_type < 0 || right_type > 2) throw arg_failure{"match mode must be 0 (nearest_sm), 1 (nearest_eq) or 2 (furthest_eq)"};
        PSAC_CUDA(cudaSetDevice(e->device));
        size_t* tot = &e->device_bytes;
        e->tb[1].reserve(n * (size_t)val_bytes, tot);
        e->tb[2].reserve(n * sizeof(u64), tot);
        e->tb[3].reserve(n * sizeof(u64), tot);
        PSAC_CUDA(cudaMemcpyAsync(e->tb[1].p, vals, n * (size_t)val_bytes, cudaMemcpyHostToDevice, e->stream));
        if (val_bytes == 4)
            ansv_device<u32>(e, e->tb[1].as<u32>(), n, left_type, right_type, nonsv, e->tb[2].as<u64>(), e->tb[3].as<u64>());
        else
            ansv_device<u64>(e, e->tb[1].as<u64>(), n, left_type, right_type, nonsv, e->tb[2].as<u64>(), e->tb[3].as<u64>());
        PSAC_CUDA(cudaMemcpyAsync(left, e->tb[2].p, n * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaMemcpyAsync(right, e->tb[3].p, n * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}

int psacb200_suffix_tree(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, const void* sa, const void* lcp, uint64_t* nodes,
                         size_t nodes_len) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    return guarded([&]() -> int {
        if (!text || !sa || !lcp || !nodes) throw arg_failure{"null argument"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        PSAC_CUDA(cudaSetDevice(e->device));
        cudaStream_t st = e->stream;
        size_t* tot = &e->device_bytes;
        e->text.reserve(n + 64, tot);
        e->small.reserve(psacb200_engine::small_bytes(), tot);
        PSAC_CUDA(cudaMemcpyAsync(e->text.p, text, n, cudaMemcpyHostToDevice, st));
        // alphabet (reference codes)
        PSAC_CUDA(cudaMemsetAsync(e->byte_hist(), 0, 256 * sizeof(u64), st));
        byte_hist_kernel<<<grid_for(e, n / 16 + 1, 512, 4), 512, 0, st>>>(e->text.as<u8>(), n, e->byte_hist());
        e->launches += 1;
        PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 16, e->byte_hist(), 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        PSAC_CUDA(cudaStreamSynchronize(st));
        Alphabet alpha;
        alphabet_from_hist(e->h_pinned + 16, alpha);
        const size_t width = (size_t)alpha.sigma + 1;
        if (nodes_len < width * n) throw arg_failure{"nodes buffer too small: (sigma + 1) * n entries are needed"};
        e->tb[1].reserve(n * (size_t)index_bytes, tot);  // LCP
        e->tb[2].reserve(n * sizeof(u64), tot);          // left ANSV
        e->tb[3].reserve(n * sizeof(u64), tot);          // right ANSV
        e->tb[4].reserve(n * (size_t)index_bytes, tot);  // SA
        e->tb[5].reserve(width * n * sizeof(u64), tot);  // child table
        PSAC_CUDA(cudaMemcpyAsync(e->tb[1].p, lcp, n * (size_t)index_bytes, cudaMemcpyHostToDevice, st));
        PSAC_CUDA(cudaMemcpyAsync(e->tb[4].p, sa, n * (size_t)index_bytes, cudaMemcpyHostToDevice, st));
        PSAC_CUDA(cudaMemsetAsync(e->tb[5].p, 0, width * n * sizeof(u64), st));
        // ansv<index_t, furthest_eq, nearest_sm>(LCP)  (suffix_tree.hpp:62)
        if (index_bytes == 4)
            ansv_device<u32>(e, e->tb[1].as<u32>(), n, 2, 0, ANSV_NONE, e->tb[2].as<u64>(), e->tb[3].as<u64>());
        else
            ansv_device<u64>(e, e->tb[1].as<u64>(), n, 2, 0, ANSV_NONE, e->tb[2].as<u64>(), e->tb[3].as<u64>());
        TreeArgs A{};
        A.sa = e->tb[4].p;
        A.lcp = e->tb[1].p;
        A.text = e->text.as<u8>();
        A.n = n;
        A.left = e->tb[2].as<u64>();
        A.right = e->tb[3].as<u64>();
        A.sigma = alpha.sigma;
        memcpy(A.lut, alpha.lut, 256);
        A.nodes = e->tb[5].as<u64>();
        if (index_bytes == 4)
            suffix_tree_kernel<u32><<<grid_for(e, n, 256, 8), 256, 0, st>>>(A);
        else
            suffix_tree_kernel<u64><<<grid_for(e, n, 256, 8), 256, 0, st>>>(A);
        e->launches += 1;
        PSAC_CUDA(cudaGetLastError());
        PSAC_CUDA(cudaMemcpyAsync(nodes, e->tb[5].p, width * n * sizeof(u64), cudaMemcpyDeviceToHost, st));
        PSAC_CUDA(cudaStreamSynchronize(st));
        return PSACB200_OK;
    });
}

int psacb200_ansv_device(psacb200_engine* e, const void* d_vals, size_t n, int val_bytes, int left_type, int right_type, uint64_t nonsv, uint64_t* d_left,
                         uint64_t* d_right) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    return guarded([&]() -> int {
        if (!d_vals || !d_left || !d_right) throw arg_failure{"null argument"};
        if (val_bytes != 4 && val_bytes != 8) throw arg_failure{"val_bytes must be 4 or 8"};
        if (left_type < 0 || left_type > 2 || right_type < 0 || right_type > 2) throw arg_failure{"match mode must be 0 (nearest_sm), 1 (nearest_eq) or 2 (furthest_eq)"};
        PSAC_CUDA(cudaSetDevice(e->device));
        e->tr_n = 0;
        e->mark("begin");
        if (val_bytes == 4)
            ansv_device<u32>(e, reinterpret_cast<const u32*>(d_vals), n, left_type, right_type, nonsv, d_left, d_right);
        else
            ansv_device<u64>(e, reinterpret_cast<const u64*>(d_vals), n, left_type, right_type, nonsv, d_left, d_right);
        e->mark("ansv");
        PSAC_CUDA(cudaStreamSynchronize(e->stream));
        return PSACB200_OK;
    });
}

int psacb200_ansv_sharded(psacb200_engine* e, const void* d_vals_local, size_t n_local, size_t n_global, int val_bytes, int left_type, int right_type,
                          uint64_t nonsv, uint64_t* d_left_local, uint64_t* d_right_local) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (!e->nccl_comm) throw arg_failure{"psacb200_comm_init has not been called on this engine"};
        if (n_local && (!d_vals_local || !d_left_local || !d_right_local)) throw arg_failure{"null argument"};
        if (val_bytes != 4 && val_bytes != 8) throw arg_failure{"val_bytes must be 4 or 8"};
        if (left_type < 0 || left_type > 2 || right_type < 0 || right_type > 2) throw arg_failure{"match mode must be 0 (nearest_sm), 1 (nearest_eq) or 2 (furthest_eq)"};
        PSAC_CUDA(cudaSetDevice(e->device));
        if (n_global == 0) return PSACB200_OK;
        ShardComm C{reinterpret_cast<ncclComm_t>(e->nccl_comm), e->shard_rank, e->shard_world};
        e->tr_n = 0;
        e->mark("begin");
        if (val_bytes == 4)
            ansv_sharded_core<u32>(e, C, reinterpret_cast<const u32*>(d_vals_local), n_local, n_global, left_type, right_type, nonsv, d_left_local, d_right_local);
        else
            ansv_sharded_core<u64>(e, C, reinterpret_cast<const u64*>(d_vals_local), n_local, n_global, left_type, right_type, nonsv, d_left_local, d_right_local);
        return PSACB200_OK;
    });
}

int psacb200_suffix_tree_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, const void* d_sa, const void* d_lcp, uint64_t* d_nodes,
                                size_t nodes_len, uint32_t* sigma_out) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    return guarded([&]() -> int {
        if (!d_text || !d_sa || !d_lcp || !d_nodes) throw arg_failure{"null argument"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        PSAC_CUDA(cudaSetDevice(e->device));
        if (index_bytes == 4)
            suffix_tree_core<u32>(e, nullptr, d_text, n, n, reinterpret_cast<const u32*>(d_sa), reinterpret_cast<const u32*>(d_lcp), d_nodes, nodes_len, sigma_out);
        else
            suffix_tree_core<u64>(e, nullptr, d_text, n, n, reinterpret_cast<const u64*>(d_sa), reinterpret_cast<const u64*>(d_lcp), d_nodes, nodes_len, sigma_out);
        return PSACB200_OK;
    });
}

int psacb200_suffix_tree_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, const void* d_sa_local,
                                 const void* d_lcp_local, uint64_t* d_nodes_local, size_t nodes_len, uint32_t* sigma_out) {
    if (!e) {
        set_last_error("null engine");
        return PSACB200_ERR_ARG;
    }
    return guarded([&]() -> int {
        if (!e->nccl_comm) throw arg_failure{"psacb200_comm_init has not been called on this engine"};
        if (n_local && (!d_text_local || !d_sa_local || !d_lcp_local || !d_nodes_local)) throw arg_failure{"null argument"};
        if (index_bytes != 4 && index_bytes != 8) throw arg_failure{"index_bytes must be 4 or 8"};
        PSAC_CUDA(cudaSetDevice(e->device));
        if (n_global == 0) return PSACB200_OK;
        ShardComm C{reinterpret_cast<ncclComm_t>(e->nccl_comm), e->shard_rank, e->shard_world};
        if (index_bytes == 4)
            suffix_tree_core<u32>(e, &C, d_text_local, n_local, n_global, reinterpret_cast<const u32*>(d_sa_local), reinterpret_cast<const u32*>(d_lcp_local),
                                  d_nodes_local, nodes_len, sigma_out);
        else
            suffix_tree_core<u64>(e, &C, d_text_local, n_local, n_global, reinterpret_cast<const u64*>(d_sa_local), reinterpret_cast<const u64*>(d_lcp_local),
                                  d_nodes_local, nodes_len, sigma_out);
        return PSACB200_OK;
    });
}

}  // extern "C"

// ================================================================================================ several GPUs behind ONE host call
// psacb200_multi: the reference-facing object holds the whole arrays (p = 1 from the caller's point of view, SURVEY.md
// section 8b) while the GPUs of the box are sharded inside the engine: one engine + one host thread per GPU, the same
// sharded construction (sharded.cuh) over an NCCL communicator of the engines, peer memory between them through plain
// peer access (one process: no IPC).
struct psacb200_multi {
    std::vector<psacb200_engine*> eng;
    psacb200_stats stats;
};

namespace {
template <typename F>
int run_on_all(psacb200_multi* m, F&& f) {
    const int p = (int)m->eng.size();
    std::vector<int> rc(p, PSACB200_OK);
    std::vector<std::string> err(p);
    std::vector<std::thread> th;
    for (int r = 0; r < p; ++r)
        th.emplace_back([&, r]() {
            rc[r] = f(r);
            if (rc[r] != PSACB200_OK) err[r] = psacb200_last_error();  // (the message is thread-local)
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < p; ++r)
        if (rc[r] != PSACB200_OK) {
            set_last_error("GPU " + std::to_string(r) + ": " + err[r]);
            return rc[r];
        }
    return PSACB200_OK;
}
}  // namespace

extern "C" {

int psacb200_multi_create(int n_gpus, const int* dev_ids, psacb200_multi** out) {
    if (!out || n_gpus < 1 || n_gpus > 16) {
        set_last_error("psacb200_multi_create: 1..16 GPUs and a non-null out pointer");
        return PSACB200_ERR_ARG;
    }
    *out = nullptr;
    psacb200_multi* m = new psacb200_multi();
    memset(&m->stats, 0, sizeof(m->stats));
    for (int r = 0; r < n_gpus; ++r) {
        psacb200_engine* e = nullptr;
        const int rc = psacb200_create(dev_ids ? dev_ids[r] : r, &e);
        if (rc != PSACB200_OK) {
            psacb200_multi_destroy(m);
            return rc;
        }
        m->eng.push_back(e);
    }
    if (n_gpus > 1) {
        uint8_t id[128];
        int rc = psacb200_comm_unique_id(id);
        if (rc == PSACB200_OK) rc = run_on_all(m, [&](int r) { return psacb200_comm_init(m->eng[r], id, r, n_gpus); });
        if (rc != PSACB200_OK) {
            psacb200_multi_destroy(m);
            return rc;
        }
    }
    *out = m;
    return PSACB200_OK;
}

void psacb200_multi_destroy(psacb200_multi* m) {
    if (!m) return;
    if (m->eng.size() > 1) run_on_all(m, [&](int r) { return m->eng[r]->nccl_comm ? psacb200_comm_finalize(m->eng[r]) : PSACB200_OK; });
    for (psacb200_engine* e : m->eng) psacb200_destroy(e);
    delete m;
}

int psacb200_multi_gpus(const psacb200_multi* m) { return m ? (int)m->eng.size() : 0; }

psacb200_engine* psacb200_multi_engine(const psacb200_multi* m, int r) { return (m && r >= 0 && r < (int)m->eng.size()) ? m->eng[r] : nullptr; }

int psacb200_multi_get_stats(const psacb200_multi* m, psacb200_stats* out) {
    if (!m || !out || m->eng.empty()) {
        set_last_error("null argument");
        return PSACB200_ERR_ARG;
    }
    return psacb200_get_stats(m->eng[0], out);
}

int psacb200_multi_construct(psacb200_multi* m, const uint8_t* text, size_t n, int index_bytes, unsigned flags, unsigned k, void* sa_out, void* isa_out,
                             void* lcp_out) {
    if (!m || m->eng.empty()) {
        set_last_error("null handle");
        return PSACB200_ERR_ARG;
    }
    const int p = (int)m->eng.size();
    if (p == 1) return psacb200_construct(m->eng[0], text, n, index_bytes, flags, k, sa_out, isa_out, lcp_out);
    if (index_bytes != 4 && index_bytes != 8) {
        set_last_error("index_bytes must be 4 or 8");
        return PSACB200_ERR_ARG;
    }
    if (n > 0 && (!text || !sa_out)) {
        set_last_error("null text / sa_out");
        return PSACB200_ERR_ARG;
    }
    if ((flags & PSACB200_LCP) && n > 0 && !lcp_out) {
        set_last_error("PSACB200_LCP set but lcp_out is null");
        return PSACB200_ERR_ARG;
    }
    if (n == 0) return PSACB200_OK;
    const bool want_lcp = (flags & PSACB200_LCP) != 0;
    return run_on_all(m, [&](int r) -> int {
        psacb200_engine* e = m->eng[r];
        return guarded([&]() -> int {
            PSAC_CUDA(cudaSetDevice(e->device));
            const BlkDist blk(n, p);
            const u64 lo = blk.start(r), nl = blk.size(r);
            size_t* tot = &e->device_bytes;
            const size_t ob = (nl + 16) * (size_t)index_bytes;
            e->tb[2].reserve(nl + 64, tot);
            e->tb[3].reserve(ob, tot);
            e->tb[4].reserve(ob, tot);
            if (want_lcp) e->tb[5].reserve(ob, tot);
            if (nl) PSAC_CUDA(cudaMemcpyAsync(e->tb[2].p, text + lo, nl, cudaMemcpyHostToDevice, e->stream));
            PSAC_CUDA(cudaStreamSynchronize(e->stream));
            const int rc = psacb200_construct_sharded(e, e->tb[2].as<u8>(), nl, n, index_bytes, flags, k, e->tb[3].p, e->tb[4].p, want_lcp ? e->tb[5].p : nullptr);
            if (rc != PSACB200_OK) return rc;
            const size_t off = (size_t)lo * index_bytes, bytes = (size_t)nl * index_bytes;
            if (bytes) {
                PSAC_CUDA(cudaMemcpyAsync(reinterpret_cast<u8*>(sa_out) + off, e->tb[3].p, bytes, cudaMemcpyDeviceToHost, e->stream));
                if (isa_out) PSAC_CUDA(cudaMemcpyAsync(reinterpret_cast<u8*>(isa_out) + off, e->tb[4].p, bytes, cudaMemcpyDeviceToHost, e->stream));
                if (want_lcp) PSAC_CUDA(cudaMemcpyAsync(reinterpret_cast<u8*>(lcp_out) + off, e->tb[5].p, bytes, cudaMemcpyDeviceToHost, e->stream));
            }
            PSAC_CUDA(cudaStreamSynchronize(e->stream));
            return PSACB200_OK;
        });
    });
}

}  // extern "C"
