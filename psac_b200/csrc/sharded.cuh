// psac-b200: suffix-array construction sharded over the GPUs of one NVLink / NVSwitch box (one process per GPU).
//
// Reference behaviour restated: text, SA, ISA and LCP are block-distributed over p ranks exactly
// like mxx::blk_dist (ext/mxx/include/mxx/partition.hpp:283-331: the first n % p ranks hold one element more); the
// reference moves (B1, B2, index) tuples through a distributed sample sort (include/idxsort.hpp:22-83 ->
// ext/mxx/include/mxx/samplesort.hpp:292-444) and (index, rank) pairs through bulk_permute_inplace
// (include/bulk_permute.hpp:14-73) every round.
//
// B200-first design, not that schedule:
//   * the PACKED text is replicated on every GPU (2 bits per DNA character: 8 Gi characters = 2 GiB of 180 GB HBM; one
//     all-gather over NVLink).  Every rank can then form the sort key of ANY suffix locally, so the sample-sort tuple
//     exchange disappears: rank r simply selects, from the whole text, the suffixes whose key prefix falls in its
//     splitter range and sorts them locally.  Splitters are exact bin boundaries of a global 14-bit key-prefix
//     histogram (all-gathered, 128 KiB per rank), so equal keys never straddle two ranks and round 0 needs no carry.
//   * the only bulk exchange is the SA -> ISA permute: (suffix, bucket id) pairs partitioned by owner with one radix
//     pass and moved with grouped ncclSend / ncclRecv (all-to-all-v on NVSwitch), then scattered locally; plus the
//     final re-balancing of SA / LCP from key-range ownership to exact blocks (contiguous ranges to <= 2 neighbours).
//   * later rounds touch only the unresolved suffixes (n / 2^10 of random text): their lists are all-gathered and the
//     rounds run replicated on every GPU; the rank look-ups ISA[s + h] are answered by the owning shard and combined with
//     one ncclAllReduce per round; every shard applies the updates that fall in its own SA / ISA windows.
// Limits of this round (checked, reported as errors): the unresolved set must fit one GPU (replicated rounds) and the
// key-prefix histogram must balance within the 2x capacity; inputs too small to shard run replicated on every GPU.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

namespace {

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;  // optional (NCCL >= 2.18)

    void load() {
        if (handle) return;
        // prefer the copy torch already loaded into this process, then the loader's search path
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW);
        if (!handle) throw std::string("cannot load libnccl.so.2: ") + dlerror();
#define PSAC_NCCL_SYM(field, name)                                                \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));               \
    if (!field) throw std::string("libnccl is missing symbol ") + name
        PSAC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        PSAC_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        PSAC_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        PSAC_NCCL_SYM(AllReduce, "ncclAllReduce");
        PSAC_NCCL_SYM(AllGather, "ncclAllGather");
        PSAC_NCCL_SYM(Broadcast, "ncclBroadcast");
        PSAC_NCCL_SYM(Send, "ncclSend");
        PSAC_NCCL_SYM(Recv, "ncclRecv");
        PSAC_NCCL_SYM(GroupStart, "ncclGroupStart");
        PSAC_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        PSAC_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PSAC_NCCL_SYM
        CommSplit = reinterpret_cast<decltype(CommSplit)>(dlsym(handle, "ncclCommSplit"));
    }
};
NcclApi g_nccl;

#define PSAC_NCCL(call)                                                                                          \
    do {                                                                                                         \
        ncclResult_t _r = (call);                                                                                \
        if (_r != ncclSuccess) throw std::string("NCCL error: ") + g_nccl.GetErrorString(_r) + " in " + #call;   \
    } while (0)

// ------------------------------------------------------------------------------------------------ host-side plans
// mxx::blk_dist (reference ext/mxx/include/mxx/partition.hpp:283-331)
struct BlkDist {
    u64 n = 0;
    int p = 1;
    u64 base = 0, rem = 0;
    BlkDist() {}
    BlkDist(u64 n_, int p_) : n(n_), p(p_), base(n_ / (u64)p_), rem(n_ % (u64)p_) {}
    u64 size(int r) const { return base + ((u64)r < rem ? 1 : 0); }
    u64 start(int r) const { return base * (u64)r + std::min<u64>((u64)r, rem); }
    int owner(u64 g) const {
        const u64 cut = rem * (base + 1);
        return g < cut ? (int)(g / (base + 1)) : (int)(rem + (g - cut) / base);
    }
};

// Splitters over a histogram of key-prefix bins: rank r owns bins [first[r], first[r+1]); greedy cut at the first bin
// boundary where the running count reaches r * n / p, so that no bin (= no set of equal keys) is split.
static void choose_splitters(const u64* hist, size_t nbins, u64 n, int p, std::vector<size_t>& first, std::vector<u64>& count) {
    first.assign(p + 1, nbins);
    count.assign(p, 0);
    first[0] = 0;
    u64 run = 0;
    int r = 1;
    for (size_t b = 0; b < nbins; ++b) {
        while (r < p && run >= (n / (u64)p) * (u64)r + std::min<u64>((u64)r, n % (u64)p)) first[r++] = b;
        run += hist[b];
    }
    while (r < p) first[r++] = nbins;
    for (int q = 0; q < p; ++q)
        for (size_t b = first[q]; b < first[q + 1]; ++b) count[q] += hist[b];
}

// ------------------------------------------------------------------------------------------------ kernels
constexpr int PREFIX_BITS_MAX = 14;

// histogram of the key-prefix bins of the suffixes [g0, g0 + cnt)
__global__ void __launch_bounds__(512) prefix_hist_kernel(const u64* __restrict__ stream, u64 g0, u64 cnt, int lbits, int pbits, u64* __restrict__ hist) {
    extern __shared__ u32 sh_hist[];
    const u32 nb = 1u << pbits;
    for (u32 e = threadIdx.x; e < nb; e += blockDim.x) sh_hist[e] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (u64)gridDim.x * blockDim.x)
        atomicAdd(&sh_hist[(u32)stream_bits(stream, (g0 + i) * (u64)lbits, pbits)], 1u);
    __syncthreads();
    for (u32 e = threadIdx.x; e < nb; e += blockDim.x) {
        const u32 c = sh_hist[e];
        if (c) atomicAdd((unsigned long long*)&hist[e], (unsigned long long)c);
    }
}

// Selection of the suffixes whose key-prefix bin lies in [bin_lo, bin_hi) out of ALL suffixes [0, n_main) of the
// replicated text: (key, suffix) pairs appended after the tails (select_tails_kernel: the suffixes that run past the end
// of the text must come first, see TextSrc in radix_sort.cuh).  One thread walks the characters of one 64-bit stream word
// with a two-word window; a CTA reserves its output range with one atomicAdd, so the order of the pairs between CTAs
// is arbitrary -- equal keys of non-tail suffixes form an unresolved bucket whose internal order is settled by the later
// rounds, and the final arrays are unique.
constexpr int SEL_THREADS = 256;

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const u64* __restrict__ stream, u64 n_main, int lbits, int kbits, int pbits, u32 bin_lo,
                                                             u32 bin_hi, u64* __restrict__ cursor, u64* __restrict__ keys_out, u64* __restrict__ suf_out) {
    __shared__ u32 s_w[SEL_THREADS / 32];
    __shared__ u64 s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cpw = 64 / lbits;
    const u64 w = (u64)blockIdx.x * SEL_THREADS + tid;
    const u64 g0 = w * (u64)cpw;
    u64 mask = 0, hi = 0, lo = 0;
    if (g0 < n_main) {
        hi = __ldg(stream + w);
        lo = __ldg(stream + w + 1);
        const int cnt = (n_main - g0 < (u64)cpw) ? (int)(n_main - g0) : cpw;
        for (int c = 0; c < cnt; ++c) {
            const int o = c * lbits;
            const u64 v = o ? ((hi << o) | (lo >> (64 - o))) : hi;
            const u32 bin = (u32)(v >> (64 - pbits));
            if (bin >= bin_lo && bin < bin_hi) mask |= 1ull << c;
        }
    }
    const u32 cnt = __popcll(mask);
    const u32 incl = warp_inclusive_sum_u32(cnt);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u32 pre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SEL_THREADS / 32; ++i) {
        if (i < warp) pre += s_w[i];
        tot += s_w[i];
    }
    if (tid == 0) s_base = tot ? atomicAdd((unsigned long long*)cursor, (unsigned long long)tot) : 0;
    __syncthreads();
    // write-out transposed: in step c every lane offers character c of ITS word, the selected lanes of the warp write
    // consecutive pairs -> each store instruction covers one contiguous run instead of 32 separate sectors
    u64 wbase = s_base + (u64)pre;  // `pre` = pairs of the warps before mine; lanes of one warp share it
    const u32 lt = lanemask_lt();
    const u64 any = __reduce_or_sync(0xffffffffu, (u32)(mask != 0));
    if (any) {
        for (int c = 0; c < cpw; ++c) {
            const bool sel = (mask >> c) & 1ull;
            const u32 b = __ballot_sync(0xffffffffu, sel);
            if (sel) {
                const int sh = c * lbits;
                const u64 v = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
                const u64 dst = wbase + __popc(b & lt);
                keys_out[dst] = v >> (64 - kbits);
                suf_out[dst] = g0 + (u64)c;
            }
            wbase += __popc(b);
        }
    }
}

// the T <= 63 suffixes that run past the end of the text, shortest first; *count = how many fall in [bin_lo, bin_hi)
__global__ void __launch_bounds__(64) select_tails_kernel(const u64* __restrict__ stream, u64 n, u64 T, int lbits, int kbits, int pbits, u32 bin_lo, u32 bin_hi,
                                                          u64* __restrict__ count, u64* __restrict__ keys_out, u64* __restrict__ suf_out) {
    __shared__ u32 s_cnt[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool sel = false;
    u64 g = 0;
    if ((u64)tid < T) {
        g = n - 1 - (u64)tid;
        const u32 bin = (u32)stream_bits(stream, g * (u64)lbits, pbits);
        sel = bin >= bin_lo && bin < bin_hi;
    }
    const u32 b = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_cnt[warp] = __popc(b);
    __syncthreads();
    const u32 o = (warp ? s_cnt[0] : 0) + __popc(b & lanemask_lt());
    if (sel) {
        keys_out[o] = stream_extract(stream, g, lbits, kbits);
        suf_out[o] = g;
    }
    if (tid == 0) *count = s_cnt[0] + s_cnt[1];
}

// key source of the SA -> ISA exchange partition: digit = owning rank of the suffix index under the block distribution
template <bool PEER_>
struct OwnerSrcT {
    using Stage = u64;
    using Out = u64;
    static constexpr bool FROM_TEXT = false;
    static constexpr bool PEER = PEER_;
    const u64* __restrict__ kin;
    const u64* __restrict__ vin;
    // PEER: bin d (= owning rank d) is written straight into rank d's receive buffers over NVLink; the pointers are
    // pre-biased so that the bin-relative index the kernel computes (global bin offset + slot) lands at the right place
    u64* kpeer[16];
    u64* vpeer[16];
    BlkDiv div;
    __device__ __forceinline__ Stage load_key(size_t g) const { return ld_stream(kin + g); }
    __device__ __forceinline__ u32 digit(Stage k) const {
        u64 local;
        return div.owner(k, &local);
    }
    __device__ __forceinline__ Out out_key(Stage k) const { return k; }
    __device__ __forceinline__ u64 load_val(size_t g) const { return ld_stream(vin + g); }
    __device__ __forceinline__ u8 load_aux(size_t, Stage) const { return 0; }
};
using OwnerSrc = OwnerSrcT<false>;
using OwnerPeerSrc = OwnerSrcT<true>;

// Packed form of the same exchange: one 64-bit word per suffix instead of a (suffix, bucket) pair, halving the bytes that
// cross NVLink and the bytes of the two receiver-side passes:
//     [ index of the suffix inside its owner's block | rank field | bucket id relative to the sender's first SA position ]
// The rank field holds the DESTINATION while the word is staged (it is the digit of the partition) and is replaced by the
// SOURCE rank when the word is written out, so the receiver can turn the relative bucket id back into a global one.
template <bool PEER_>
struct OwnerPackSrcT {
    using Stage = u64;
    using Out = u64;
    static constexpr bool FROM_TEXT = false;
    static constexpr bool PEER = PEER_;
    const u64* __restrict__ kin;  // suffix (global index)
    const u64* __restrict__ vin;  // bucket id (global SA position)
    u64* kpeer[16];
    u64* vpeer[16];               // unused (keys only)
    BlkDiv div;                   // block distribution
    u64 pos_base;                 // first SA position of the sender
    int rank_shift, idx_shift;    // bit positions of the rank field and of the block-local index
    u32 rank_mask;
    u64 me;
    __device__ __forceinline__ Stage load_key(size_t g) const {
        const u64 s = ld_stream(kin + g), b = ld_stream(vin + g);
        u64 local;
        const u64 owner = div.owner(s, &local);
        return (local << idx_shift) | (owner << rank_shift) | (b - pos_base);
    }
    __device__ __forceinline__ u32 digit(Stage k) const { return (u32)(k >> rank_shift) & rank_mask; }
    __device__ __forceinline__ Out out_key(Stage k) const { return (k & ~((u64)rank_mask << rank_shift)) | (me << rank_shift); }
    __device__ __forceinline__ NoVal load_val(size_t) const { return NoVal(); }
    __device__ __forceinline__ u8 load_aux(size_t, Stage) const { return 0; }
};

// receiver side of the packed exchange: ISA[local index] = first SA position of the source rank + relative bucket id
struct PackedScatterArgs {
    const u64* words;
    u64* isa;
    u64 n;
    int rank_shift, idx_shift;
    u32 rank_mask;
    u64 rel_mask;
    u64 off_key[16];
};
__global__ void __launch_bounds__(256) isa_scatter_packed_kernel(PackedScatterArgs A) {
    constexpr int PER = 16;
    const u64 base = (u64)blockIdx.x * (256 * PER);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const u64 q = base + (u64)i * 256 + threadIdx.x;
        if (q < A.n) {
            const u64 w = ld_stream(A.words + q);
            A.isa[w >> A.idx_shift] = A.off_key[(w >> A.rank_shift) & A.rank_mask] + (w & A.rel_mask);
        }
    }
}

// rank look-ups of one replicated round: ans[j] = ISA[suf[j] + h] + 1 if this shard owns that entry, else 0
__global__ void __launch_bounds__(256) isa_answer_kernel(const u64* __restrict__ suf, u64 m, u64 h, u64 n, const u64* __restrict__ isa, u64 isa_lo,
                                                         u64 isa_hi, u64* __restrict__ ans) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) {
        const u64 g = suf[j] + h;
        ans[j] = (g < n && g >= isa_lo && g < isa_hi) ? isa[g - isa_lo] + 1 : 0;
    }
}

__global__ void __launch_bounds__(32) last_pair_kernel(const u64* __restrict__ keys, const u64* __restrict__ sufs, u64 cnt, u64* __restrict__ out) {
    if (threadIdx.x == 0) {
        out[0] = cnt ? keys[cnt - 1] : 0;
        out[1] = cnt ? sufs[cnt - 1] : 0;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ sharded engine state
struct ShardComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

namespace {

// Peer-visible arena: ONE cudaMalloc'd region per rank, mapped into every peer (CUDA IPC between processes, plain peer
// access between the engines of one process), so that kernels can load and store straight in their peers' HBM over
// NVLink: the exchange steps are fused into the kernels that produce / consume the data.  Everything a peer may touch
// (exchange receive buffers, the ISA block during the later rounds and the checker) is sub-allocated from it at offsets
// that are identical on all ranks.  The region is never freed while a peer may hold a mapping: growing it is a
// collective step (every rank closes its mappings, barrier, free + malloc, handles exchanged again), decided from
// all-reduced sizes so that all ranks take it together; psacb200_comm_finalize releases it the same way.
struct PeerArena {
    u8* base = nullptr;
    size_t bytes = 0;
    u8* peer[16];
    bool ipc_open[16];
    int world = 0;
    bool usable = false;
    PeerArena() {
        memset(peer, 0, sizeof(peer));
        memset(ipc_open, 0, sizeof(ipc_open));
    }
};

// host-visible barrier: a one-word all-reduce, then the stream is drained
void host_barrier(psacb200_engine* e, const ShardComm& C) {
    u64* d = e->shard_meta() + 57;
    PSAC_NCCL(g_nccl.AllReduce(d, d, 1, ncclUint64, ncclSum, C.comm, e->stream));
    PSAC_CUDA(cudaStreamSynchronize(e->stream));
}

void arena_close_peers(PeerArena& A) {
    for (int r = 0; r < 16; ++r) {
        if (A.ipc_open[r]) cudaIpcCloseMemHandle(A.peer[r]);
        A.ipc_open[r] = false;
        A.peer[r] = nullptr;
    }
    A.usable = false;
}

// Collective.  Makes sure every rank owns an arena of at least `need` bytes that all peers have mapped.  Returns false --
// on every rank -- when peer mapping is unavailable (then the callers use their NCCL paths).
bool arena_ensure(psacb200_engine* e, const ShardComm& C, size_t need) {
    if (e->peer_map == nullptr) return false;  // PSACB200_NO_PEER
    PeerArena& A = *reinterpret_cast<PeerArena*>(e->peer_map);
    cudaStream_t st = e->stream;
    const int p = C.world, me = C.rank;
    // agree on max(need) and min(current size)
    u64* d_w = e->shard_meta() + 58;  // 2 words
    e->h_pinned[40] = (u64)need;
    e->h_pinned[41] = ~(u64)((A.usable && A.world == p) ? A.bytes : 0);
    PSAC_CUDA(cudaMemcpyAsync(d_w, e->h_pinned + 40, 2 * sizeof(u64), cudaMemcpyHostToDevice, st));
    PSAC_NCCL(g_nccl.AllReduce(d_w, d_w, 2, ncclUint64, ncclMax, C.comm, st));
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 40, d_w, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    const u64 need_all = e->h_pinned[40], have_min = ~e->h_pinned[41];
    if (have_min >= need_all && have_min > 0) return true;
    // ---- (re)allocate together
    arena_close_peers(A);
    host_barrier(e, C);  // nobody maps my old region any more
    if (A.base) {
        cudaFree(A.base);
        e->device_bytes -= A.bytes;
        A.base = nullptr;
        A.bytes = 0;
    }
    const size_t bytes = align_up((size_t)need_all + (size_t)need_all / 16, (size_t)2 << 20);
    int ok = 1;
    struct Info {
        cudaIpcMemHandle_t handle;
        u64 pid, ptr;
        int dev, ok;
        u8 pad[128 - 64 - 16 - 8];
    } mine;
    static_assert(sizeof(Info) == 128, "arena info record");
    memset(&mine, 0, sizeof(mine));
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    } else {
        A.base = reinterpret_cast<u8*>(ptr);
        A.bytes = bytes;
        e->device_bytes += bytes;
        if (cudaIpcGetMemHandle(&mine.handle, ptr) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
    }
    mine.pid = (u64)getpid();
    mine.ptr = (u64)(uintptr_t)ptr;
    mine.dev = e->device;
    mine.ok = ok;
    e->scratch.reserve((size_t)p * 128 + 64, &e->device_bytes);
    u8* d_h = e->scratch.as<u8>();
    PSAC_CUDA(cudaMemcpyAsync(d_h + (size_t)me * 128, &mine, 128, cudaMemcpyHostToDevice, st));
    PSAC_NCCL(g_nccl.AllGather(d_h + (size_t)me * 128, d_h, 128, ncclUint8, C.comm, st));
    std::vector<Info> all(p);
    PSAC_CUDA(cudaMemcpyAsync(all.data(), d_h, (size_t)p * 128, cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < p; ++r) ok &= all[r].ok;
    for (int r = 0; r < p && ok; ++r) {
        if (r == me) {
            A.peer[r] = A.base;
        } else if (all[r].pid == mine.pid) {
            // another engine of this process: plain peer access
            cudaError_t pe = all[r].dev == e->device ? cudaSuccess : cudaDeviceEnablePeerAccess(all[r].dev, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
            cudaGetLastError();
            A.peer[r] = reinterpret_cast<u8*>((uintptr_t)all[r].ptr);
        } else {
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            } else {
                A.peer[r] = reinterpret_cast<u8*>(q);
                A.ipc_open[r] = true;
            }
        }
    }
    // agree: one failing rank sends everybody to the NCCL paths
    e->h_pinned[40] = (u64)ok;
    PSAC_CUDA(cudaMemcpyAsync(d_w, e->h_pinned + 40, sizeof(u64), cudaMemcpyHostToDevice, st));
    PSAC_NCCL(g_nccl.AllReduce(d_w, d_w, 1, ncclUint64, ncclMin, C.comm, st));
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 40, d_w, sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    A.world = p;
    A.usable = e->h_pinned[40] != 0;
    if (!A.usable) arena_close_peers(A);
    return A.usable;
}

// Collective release (psacb200_comm_finalize): mappings first, then the regions.
void arena_release(psacb200_engine* e, const ShardComm* C) {
    if (e->peer_map == nullptr) return;
    PeerArena& A = *reinterpret_cast<PeerArena*>(e->peer_map);
    cudaStreamSynchronize(e->stream);
    arena_close_peers(A);
    if (C != nullptr && C->comm != nullptr) host_barrier(e, *C);
    if (A.base) {
        cudaFree(A.base);
        e->device_bytes -= A.bytes;
    }
    A.base = nullptr;
    A.bytes = 0;
}

// stream-ordered barrier over the ranks (a one-word all-reduce)
void rank_barrier(psacb200_engine* e, const ShardComm& C) {
    u64* d = e->shard_meta() + 57;
    PSAC_NCCL(g_nccl.AllReduce(d, d, 1, ncclUint64, ncclSum, C.comm, e->stream));
}

// all-to-all-v of `elt`-byte elements on the engine's stream (counts / displacements in elements)
void all_to_all_v(psacb200_engine* e, const ShardComm& C, const void* send, const std::vector<u64>& scount, const std::vector<u64>& sdispl, void* recv,
                  const std::vector<u64>& rcount, const std::vector<u64>& rdispl, size_t elt) {
    const char* s = reinterpret_cast<const char*>(send);
    char* r = reinterpret_cast<char*>(recv);
    if (scount[C.rank])
        PSAC_CUDA(cudaMemcpyAsync(r + rdispl[C.rank] * elt, s + sdispl[C.rank] * elt, scount[C.rank] * elt, cudaMemcpyDeviceToDevice, e->stream));
    PSAC_NCCL(g_nccl.GroupStart());
    for (int peer = 0; peer < C.world; ++peer) {
        if (peer == C.rank) continue;
        if (scount[peer]) PSAC_NCCL(g_nccl.Send(s + sdispl[peer] * elt, scount[peer] * elt, ncclUint8, peer, C.comm, e->stream));
        if (rcount[peer]) PSAC_NCCL(g_nccl.Recv(r + rdispl[peer] * elt, rcount[peer] * elt, ncclUint8, peer, C.comm, e->stream));
    }
    PSAC_NCCL(g_nccl.GroupEnd());
}

// every rank contributes count[rank] elements at displ[rank] of the same replicated buffer (in place)
void all_gather_v(psacb200_engine* e, const ShardComm& C, void* buf, const std::vector<u64>& count, const std::vector<u64>& displ, size_t elt) {
    char* b = reinterpret_cast<char*>(buf);
    PSAC_NCCL(g_nccl.GroupStart());
    for (int root = 0; root < C.world; ++root)
        if (count[root]) PSAC_NCCL(g_nccl.Broadcast(b + displ[root] * elt, b + displ[root] * elt, count[root] * elt, ncclUint8, root, C.comm, e->stream));
    PSAC_NCCL(g_nccl.GroupEnd());
}

// S1 alphabet: local byte histogram, summed over the ranks (reference alphabet.hpp:94-100 allreduce);
// S2 the packed text, replicated on every rank (e->packed)
void prepare_text_sharded(psacb200_engine* e, const ShardComm& C, const u8* d_text_local, u64 n_local, u64 n, Alphabet& alpha) {
    const int p = C.world, me = C.rank;
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    psacb200_stats& S = e->stats;
    const BlkDist blk(n, p);
    e->small.reserve(psacb200_engine::small_bytes(), tot);
    e->begin(PH_ALPHABET);
    PSAC_CUDA(cudaMemsetAsync(e->byte_hist(), 0, 256 * sizeof(u64), st));
    if (n_local) {
        byte_hist_kernel<<<grid_for(e, n_local / 16 + 1, 512, 4), 512, 0, st>>>(d_text_local, n_local, e->byte_hist());
        e->launches += 1;
    }
    PSAC_NCCL(g_nccl.AllReduce(e->byte_hist(), e->byte_hist(), 256, ncclUint64, ncclSum, C.comm, st));
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 16, e->byte_hist(), 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    e->end(PH_ALPHABET);
    PSAC_CUDA(cudaStreamSynchronize(st));
    alphabet_from_hist(e->h_pinned + 16, alpha);
    dense_codes(e->h_pinned + 16, alpha);
    S.sigma = alpha.sigma;
    S.bits_per_char = alpha.ref_bits;
    S.pack_bits = alpha.lbits;
    const int lbits = alpha.lbits, cpw = 64 / lbits;

    e->begin(PH_PACK);
    const size_t nwords = div_up(n, (size_t)cpw) + 2;
    e->packed.reserve(nwords * sizeof(u64) + 256, tot);
    u64* stream = e->packed.as<u64>();
    const bool aligned = (n % (u64)p == 0) && ((n / (u64)p) % (u64)cpw == 0);
    if (aligned) {
        // every block starts on a stream-word boundary: pack the local block in place, all-gather the words
        const size_t wloc = n_local / cpw;
        PSAC_CUDA(cudaMemsetAsync(stream + (size_t)p * wloc, 0, (nwords - (size_t)p * wloc) * sizeof(u64), st));
        pack_text_kernel<<<grid_for(e, wloc, 256, 8), 256, 0, st>>>(d_text_local, n_local, alpha.dense, lbits, stream + (size_t)me * wloc, wloc);
        e->launches += 1;
        PSAC_NCCL(g_nccl.AllGather(stream + (size_t)me * wloc, stream, wloc, ncclUint64, C.comm, st));
    } else {
        // general block sizes: all-gather the raw characters, pack everything locally
        e->text.reserve(n + 64, tot);
        std::vector<u64> cnt(p), dsp(p);
        for (int r = 0; r < p; ++r) {
            cnt[r] = blk.size(r);
            dsp[r] = blk.start(r);
        }
        if (n_local) PSAC_CUDA(cudaMemcpyAsync(e->text.as<u8>() + dsp[me], d_text_local, n_local, cudaMemcpyDeviceToDevice, st));
        all_gather_v(e, C, e->text.p, cnt, dsp, 1);
        pack_text_kernel<<<grid_for(e, nwords, 256, 8), 256, 0, st>>>(e->text.as<u8>(), n, alpha.dense, lbits, stream, nwords);
        e->launches += 1;
    }
    PSAC_CUDA(cudaGetLastError());
    e->end(PH_PACK);
}

struct ShardedTimes {
    cudaEvent_t ev[12];
};

// The sharded construction proper.  d_text_local: this rank's block of the text (device).  Outputs: this rank's blocks
// of SA / ISA / LCP (device, index_bytes wide).
// Returns false -- identically on every rank, from data all of them hold -- when this round's sharded scheme cannot take
// the input (key prefixes too skewed to balance, or too many unresolved suffixes for the replicated rounds); the caller
// then runs the replicated construction.
bool construct_sharded_core(psacb200_engine* e, const ShardComm& C, const u8* d_text_local, u64 n_local, u64 n, int index_bytes, unsigned flags,
                            unsigned k, void* sa_out, void* isa_out, void* lcp_out) {
    const bool want_lcp = (flags & PSACB200_LCP) != 0;
    const int p = C.world, me = C.rank;
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    psacb200_stats& S = e->stats;
    const BlkDist blk(n, p);
    if (blk.size(me) != n_local) throw arg_failure{"the input text must be equally block decomposed across all ranks (reference suffix_array.hpp:226)"};
    S.internal_index_bytes = 8;
    S.sort_elt_bytes = 16;
    e->v1_stats = true;
    e->small.reserve(psacb200_engine::small_bytes(), tot);

    // ---- S1 + S2 alphabet and replicated packed text
    Alphabet alpha;
    prepare_text_sharded(e, C, d_text_local, n_local, n, alpha);
    const int lbits = alpha.lbits, cpw = 64 / lbits;
    u64* stream = e->packed.as<u64>();

    // ---- S3 key length, key-prefix histogram of the local block, all-gathered
    const unsigned Cc = choose_key_chars(n, lbits, k);
    S.key_chars = Cc;
    const int kbits = (int)Cc * lbits;
    const int pbits = std::min(PREFIX_BITS_MAX, kbits);
    const size_t nbins = (size_t)1 << pbits;
    e->begin(PH_HIST);
    e->scratch.reserve((size_t)p * nbins * sizeof(u64), tot);
    u64* d_hist = e->scratch.as<u64>();
    PSAC_CUDA(cudaMemsetAsync(d_hist + (size_t)me * nbins, 0, nbins * sizeof(u64), st));
    if (n_local) {
        PSAC_CUDA(cudaFuncSetAttribute(prefix_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(nbins * sizeof(u32))));
        prefix_hist_kernel<<<grid_for(e, n_local, 512, 2), 512, nbins * sizeof(u32), st>>>(stream, blk.start(me), n_local, lbits, pbits, d_hist + (size_t)me * nbins);
        e->launches += 1;
        PSAC_CUDA(cudaGetLastError());
    }
    PSAC_NCCL(g_nccl.AllGather(d_hist + (size_t)me * nbins, d_hist, nbins, ncclUint64, C.comm, st));
    std::vector<u64> h2d((size_t)p * nbins);
    PSAC_CUDA(cudaMemcpyAsync(h2d.data(), d_hist, h2d.size() * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    e->end(PH_HIST);
    std::vector<u64> hist(nbins, 0);
    for (int b = 0; b < p; ++b)
        for (size_t i = 0; i < nbins; ++i) hist[i] += h2d[(size_t)b * nbins + i];
    std::vector<size_t> first;
    std::vector<u64> cnt_key;  // suffixes per key range = SA positions owned after the sort
    choose_splitters(hist.data(), nbins, n, p, first, cnt_key);
    std::vector<u64> off_key(p + 1, 0);
    for (int r = 0; r < p; ++r) off_key[r + 1] = off_key[r] + cnt_key[r];
    const u64 cnt = cnt_key[me], off = off_key[me];
    const u64 cap = 2 * div_up(n, (size_t)p) + 4096;
    for (int r = 0; r < p; ++r)
        if (cnt_key[r] > cap || cnt_key[r] == 0 || cnt_key[r] >= (1ull << 32)) return false;  // too skewed for bin-boundary splitters (or beyond the 32-bit local positions of the heads kernel)
    // send counts of the ISA exchange: [text block b][key range a] = suffixes of block b whose bin lies in range a
    std::vector<u64> blk_in_range((size_t)p * p, 0);
    for (int b = 0; b < p; ++b)
        for (int a = 0; a < p; ++a)
            for (size_t i = first[a]; i < first[a + 1]; ++i) blk_in_range[(size_t)b * p + a] += h2d[(size_t)b * nbins + i];

    // ---- S5 select my key range from the whole text: (key, suffix) pairs, tails first
    const u64 T = (n < (u64)Cc - 1) ? n : (u64)Cc - 1;
    const u64 n_main = n - T;
    for (int b = 0; b < 2; ++b) {
        e->keys[b].reserve((cnt + 16) * sizeof(u64), tot);
        e->vals[b].reserve((cnt + 16) * sizeof(u64), tot);
    }
    const bool isa_inplace = index_bytes == 8 && isa_out != nullptr;  // 64-bit caller: scatter straight into the caller's ISA block
    if (!isa_inplace) e->isa.reserve((n_local + 16) * sizeof(u64), tot);
    if (want_lcp) e->lcp.reserve((cnt + 16) * sizeof(u64), tot);
    e->lookback.reserve(lookback_bytes(std::max(cnt, n_local)), tot);
    e->begin(PH_SORT);
    u64* d_cursor = e->shard_meta() + 48;  // number of pairs written so far: the tails first, then the selection appends
    select_tails_kernel<<<1, 64, 0, st>>>(stream, n, T, lbits, kbits, pbits, (u32)first[me], (u32)first[me + 1], d_cursor, e->keys[0].as<u64>(),
                                          e->vals[0].as<u64>());
    if (n_main) {
        const u64 sel_ctas = div_up(div_up(n_main, (size_t)cpw), (size_t)SEL_THREADS);
        select_kernel<<<(unsigned)sel_ctas, SEL_THREADS, 0, st>>>(stream, n_main, lbits, kbits, pbits, (u32)first[me], (u32)first[me + 1], d_cursor,
                                                                  e->keys[0].as<u64>(), e->vals[0].as<u64>());
    }
    e->launches += 2;
    PSAC_CUDA(cudaGetLastError());
    cudaEventRecord(e->ev_end[PH_PASS1], st);  // selection time is reported in the slot of "digit pass 1"

    // ---- S6 local sort by the whole key (stable: tails stay in front of equal keys)
    uint64_t sl = 0;
    RadixPlan plan_used;
    const bool alt = radix_sort_pairs<u64, u64>(e->radix_ws(), e->keys[0].as<u64>(), e->keys[1].as<u64>(), e->vals[0].as<u64>(), e->vals[1].as<u64>(), cnt, 0,
                                                kbits, st, e->sm_count, &plan_used, &sl);
    e->launches += sl;
    S.sort_passes = plan_used.npass;
    e->end(PH_SORT);
    const int x = alt ? 1 : 0, y = 1 - x;
    u64* SA = e->vals[x].as<u64>();     // SA positions [off, off + cnt)
    u64* ISA = isa_inplace ? reinterpret_cast<u64*>(isa_out) : e->isa.as<u64>();  // ISA entries of my text block
    u64* LCP = want_lcp ? e->lcp.as<u64>() : nullptr;
    const u64 text_lo = blk.start(me), text_hi = text_lo + n_local;

    // ---- S7 resolve round 0.  The element before my first one is the last element of the previous key range.
    e->begin(PH_RESOLVE);
    u64* d_last = e->shard_meta();                    // [2 * p] all-gathered {last key, last suffix}
    last_pair_kernel<<<1, 32, 0, st>>>(e->keys[x].as<u64>(), SA, cnt, d_last + 2 * me);
    PSAC_NCCL(g_nccl.AllGather(d_last + 2 * me, d_last, 2, ncclUint64, C.comm, st));
    u64* bucket = e->keys[y].as<u64>();
    const u64 ucap = std::max<u64>(unresolved_cap(cnt), 1);
    e->rp[1].reserve(ucap * sizeof(u64), tot);
    e->rh[1].reserve(ucap, tot);
    e->rv[1].reserve(ucap * sizeof(u64), tot);
    ResolveArgs R{};
    R.keys = e->keys[x].p;
    R.vals = SA;
    R.pos_in = nullptr;
    R.m = cnt;
    R.n = n;
    R.sa = SA;
    R.isa = nullptr;
    R.bucket_out = bucket;
    R.lcp = LCP;
    R.pos_out = e->rp[1].p;
    R.head_out = e->rh[1].as<u8>();
    R.suf_out = e->rv[1].p;
    R.cap = ucap;
    R.counts = e->counts();
    R.lb_max = e->lookback.as<u64>();
    R.lb_sum = R.lb_max + div_up(cnt, (size_t)RES_TILE);
    R.stream = stream;
    R.lbits = lbits;
    R.C = (int)Cc;
    R.kbits = 0;
    R.h = 0;
    R.padded_lcp = alpha.zero_code_used ? 1 : 0;
    R.pos_base = off;
    R.halo = me > 0 ? d_last + 2 * (me - 1) : nullptr;
    R.sa_lo = off;
    R.sa_hi = off + cnt;
    R.isa_lo = text_lo;
    R.isa_hi = text_hi;
    if (!alpha.zero_code_used) {
        // lean round-0 resolve (sa_kernels.cuh heads_kernel) with global positions
        TailList* tails = reinterpret_cast<TailList*>(e->tail_list());
        tail_positions_kernel<u64><<<1, 64, 0, st>>>(e->keys[x].as<u64>(), cnt, nullptr, 0, stream, n, T, lbits, kbits, pbits, (u32)first[me],
                                                     (u32)first[me + 1], tails);
        HeadsArgs H{};
        H.keys = e->keys[x].p;
        H.seg_dense = nullptr;
        H.seg_shift = 0;
        H.seg_flags = nullptr;
        H.vals = SA;
        H.m = cnt;
        H.n = n;
        H.lbits = lbits;
        H.C = (int)Cc;
        H.tails = tails;
        H.bucket_out = bucket;
        H.isa = nullptr;
        H.lcp = LCP;
        H.pos_out = R.pos_out;
        H.head_out = R.head_out;
        H.suf_out = R.suf_out;
        H.cap = R.cap;
        H.counts = R.counts;
        H.pos_base = off;
        H.halo = R.halo;
        const u64 ntiles = div_up(cnt, (size_t)HD_TILE);
        H.agg_max = e->lookback.as<u64>();
        H.agg_sum = H.agg_max + ntiles;
        PSAC_CUDA(cudaMemsetAsync(R.counts, 0, 2 * sizeof(u64), st));
        heads_kernel<u64, u64, 0><<<(unsigned)ntiles, HD_THREADS, 0, st>>>(H);
        tile_scan_kernel<<<1, 1024, 0, st>>>(H.agg_max, H.agg_sum, ntiles, R.counts);
        heads_kernel<u64, u64, 1><<<(unsigned)ntiles, HD_THREADS, 0, st>>>(H);
        e->launches += 4;
        PSAC_CUDA(cudaGetLastError());
    } else {
        launch_resolve<u64, u64>(e, true, R);
    }
    e->end(PH_RESOLVE);
    // unresolved counts of all ranks
    u64* d_m = e->shard_meta() + 32;  // [p]
    PSAC_CUDA(cudaMemcpyAsync(d_m + me, e->counts(), sizeof(u64), cudaMemcpyDeviceToDevice, st));
    PSAC_NCCL(g_nccl.AllGather(d_m + me, d_m, 1, ncclUint64, C.comm, st));
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned, d_m, (size_t)p * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    std::vector<u64> m_r(p), m_dsp(p);
    u64 M = 0;
    for (int r = 0; r < p; ++r) {
        m_r[r] = e->h_pinned[r];
        m_dsp[r] = M;
        M += m_r[r];
    }
    for (int r = 0; r < p; ++r)
        if (m_r[r] > std::max<u64>(unresolved_cap(cnt_key[r]), 1)) return false;  // some rank's list overflowed (repetitive text)
    if (M > (1ull << 28)) return false;                                          // too many for the replicated rounds
    S.unresolved_after_first = M;
    S.rounds = 1;

    // ---- S9 SA -> ISA: partition (suffix, bucket) by owning rank, all-to-all-v, scatter into my ISA block
    e->begin(PH_ISA);
    {
        RadixWorkspace ws = e->radix_ws();
        std::vector<u64> scount(p), sdispl(p), rcount(p), rdispl(p);
        u64 run = 0;
        for (int b = 0; b < p; ++b) {
            scount[b] = blk_in_range[(size_t)b * p + me];  // my key range, text block b
            sdispl[b] = run;
            run += scount[b];
        }
        if (run != cnt) throw std::string("sharded construction: exchange plan does not add up");
        u64 rrun = 0;
        for (int a = 0; a < p; ++a) {
            rcount[a] = blk_in_range[(size_t)me * p + a];
            rdispl[a] = rrun;
            rrun += rcount[a];
        }
        if (rrun != n_local) throw std::string("sharded construction: exchange plan does not cover the block");
        // receive buffers of n_local pairs
        // receive buffers of n_local pairs: in the peer-visible arena when it can be mapped, else private (NCCL path)
        const size_t rbytes = align_up((blk.size(0) + 16) * sizeof(u64), 256);
        const bool fused = arena_ensure(e, C, 2 * rbytes);
        u64* recv_suf;
        u64* recv_bkt;
        u64* peer_suf[16];
        u64* peer_bkt[16];
        if (fused) {
            PeerArena& A = *reinterpret_cast<PeerArena*>(e->peer_map);
            recv_suf = reinterpret_cast<u64*>(A.base);
            recv_bkt = reinterpret_cast<u64*>(A.base + rbytes);
            for (int r = 0; r < p; ++r) {
                peer_suf[r] = reinterpret_cast<u64*>(A.peer[r]);
                peer_bkt[r] = reinterpret_cast<u64*>(A.peer[r] + rbytes);
            }
        } else {
            e->rk[0].reserve((n_local + 16) * sizeof(u64), tot);
            e->rk[1].reserve((n_local + 16) * sizeof(u64), tot);
            recv_suf = e->rk[0].as<u64>();
            recv_bkt = e->rk[1].as<u64>();
        }
        S.reserved = fused ? 1u : 0u;  // reported as "exchange = peer stores" in the stats
        // packed exchange: [local index | rank | relative bucket] in one word, when the three fields fit 64 bits
        u64 max_cnt = 0;
        for (int r = 0; r < p; ++r) max_cnt = std::max(max_cnt, cnt_key[r]);
        const int rel_bits = (int)bits_for(max_cnt ? max_cnt - 1 : 0), rank_bits = std::max(1, (int)bits_for((u64)p - 1));
        const int idx_bits = (int)bits_for(blk.size(0) ? blk.size(0) - 1 : 0);
        const bool packed = rel_bits + rank_bits + idx_bits <= 64 && !getenv("PSACB200_NO_PACK");  // PSACB200_NO_PACK=1: (suffix, bucket) pairs
        if (packed) {
            const int rank_shift = rel_bits, idx_shift = rel_bits + rank_bits;
            u64* recv_words = recv_suf;
            if (fused) {
                OwnerPackSrcT<true> src{SA, bucket, {}, {}, BlkDiv::make(n, p), off, rank_shift,
                                        idx_shift, (1u << rank_bits) - 1u, (u64)me};
                for (int b = 0; b < p; ++b) {
                    u64 rd = 0;
                    for (int a = 0; a < me; ++a) rd += blk_in_range[(size_t)b * p + a];
                    src.kpeer[b] = peer_suf[b] + rd - sdispl[b];
                    src.vpeer[b] = nullptr;
                }
                for (int b = p; b < 16; ++b) src.kpeer[b] = src.vpeer[b] = nullptr;
                rank_barrier(e, C);
                launch_pass<OwnerPackSrcT<true>, NoVal, false>(ws, src, nullptr, nullptr, nullptr, cnt, st);
                e->launches += LAUNCHES_PER_PASS;
                rank_barrier(e, C);
            } else {
                u64* part_words = e->vals[y].as<u64>();
                OwnerPackSrcT<false> src{SA, bucket, {}, {}, BlkDiv::make(n, p), off, rank_shift,
                                         idx_shift, (1u << rank_bits) - 1u, (u64)me};
                launch_pass<OwnerPackSrcT<false>, NoVal, false>(ws, src, part_words, nullptr, nullptr, cnt, st);
                e->launches += LAUNCHES_PER_PASS;
                all_to_all_v(e, C, part_words, scount, sdispl, recv_words, rcount, rdispl, sizeof(u64));
            }
            // window partition by the top 8 bits of the local index, then the windowed scatter
            const int shift2 = idx_shift + (idx_bits > RADIX_BITS ? idx_bits - RADIX_BITS : 0);
            ArraySrc<u64, NoVal> wsrc{recv_words, nullptr, nullptr, shift2, (u32)(RADIX - 1), 0ull};
            u64* win_words = recv_bkt;  // the second receive buffer is free in the packed exchange
            launch_pass<ArraySrc<u64, NoVal>, NoVal, false>(ws, wsrc, win_words, nullptr, nullptr, n_local, st);
            PackedScatterArgs PA{};
            PA.words = win_words;
            PA.isa = ISA;
            PA.n = n_local;
            PA.rank_shift = rank_shift;
            PA.idx_shift = idx_shift;
            PA.rank_mask = (1u << rank_bits) - 1u;
            PA.rel_mask = rel_bits >= 64 ? ~0ull : ((1ull << rel_bits) - 1ull);
            for (int r = 0; r < 16; ++r) PA.off_key[r] = r < p ? off_key[r] : 0;
            isa_scatter_packed_kernel<<<(unsigned)div_up(n_local, (size_t)4096), 256, 0, st>>>(PA);
            e->launches += LAUNCHES_PER_PASS + 1;
            PSAC_CUDA(cudaGetLastError());
        } else if (fused) {
            // ONE kernel partitions by owner and stores each bin into its owner's receive buffer over NVLink (peer
            // stores): the exchange overlaps the partition tile by tile.  Rank b receives my bin at its displacement
            // for source `me`; barriers keep the receive buffers of a previous call / the next step apart.
            OwnerPeerSrc src{SA, bucket, {}, {}, BlkDiv::make(n, p)};
            for (int b = 0; b < p; ++b) {
                u64 rd = 0;  // displacement of source `me` in receiver b's buffer
                for (int a = 0; a < me; ++a) rd += blk_in_range[(size_t)b * p + a];
                src.kpeer[b] = peer_suf[b] + rd - sdispl[b];
                src.vpeer[b] = peer_bkt[b] + rd - sdispl[b];
            }
            for (int b = p; b < 16; ++b) src.kpeer[b] = src.vpeer[b] = nullptr;
            rank_barrier(e, C);
            launch_pass<OwnerPeerSrc, u64, false>(ws, src, nullptr, nullptr, nullptr, cnt, st);
            e->launches += LAUNCHES_PER_PASS;
            rank_barrier(e, C);
        } else {
            u64* part_suf = e->vals[y].as<u64>();
            u64* part_bkt = e->keys[x].as<u64>();  // the sorted keys are dead after resolve
            OwnerSrc src{SA, bucket, {}, {}, BlkDiv::make(n, p)};
            launch_pass<OwnerSrc, u64, false>(ws, src, part_suf, part_bkt, nullptr, cnt, st);
            e->launches += LAUNCHES_PER_PASS;
            all_to_all_v(e, C, part_suf, scount, sdispl, recv_suf, rcount, rdispl, sizeof(u64));
            all_to_all_v(e, C, part_bkt, scount, sdispl, recv_bkt, rcount, rdispl, sizeof(u64));
        }
        if (packed) {
            // done above
        } else if (n_local >= (1ull << 22)) {
            // as on one GPU: partition the received pairs by ISA window (top 8 bits of the index inside my block), then
            // scatter window by window so the writes stay in L2 until their sectors are complete
            const int nb2 = (int)bits_for(n_local - 1);
            const int shift2 = nb2 > RADIX_BITS ? nb2 - RADIX_BITS : 0;
            ArraySrc<u64, u64> wsrc{recv_suf, recv_bkt, nullptr, shift2, (u32)(RADIX - 1), text_lo};
            e->vals[y].reserve((n_local + 16) * sizeof(u64), tot);
            e->keys[x].reserve((n_local + 16) * sizeof(u64), tot);
            u64* win_suf = e->vals[y].as<u64>();  // the partition buffers of the send side are free again
            u64* win_bkt = e->keys[x].as<u64>();
            launch_pass<ArraySrc<u64, u64>, u64, false>(ws, wsrc, win_suf, win_bkt, nullptr, n_local, st);
            isa_scatter_kernel<u64><<<(unsigned)div_up(n_local, (size_t)4096), 256, 0, st>>>(win_suf, win_bkt, ISA - text_lo, n_local);
            e->launches += LAUNCHES_PER_PASS + 1;
        } else if (n_local) {
            isa_scatter_kernel<u64><<<(unsigned)div_up(n_local, (size_t)4096), 256, 0, st>>>(recv_suf, recv_bkt, ISA - text_lo, n_local);
            e->launches += 1;
        }
        PSAC_CUDA(cudaGetLastError());
    }
    e->end(PH_ISA);

    // ---- S8 later rounds, replicated on every rank over the all-gathered unresolved set
    if (M > 0) {
        e->begin(PH_ROUNDS);
        // gathered lists first (rp/rh/rv[0]); the local lists live in rp/rh/rv[1], which may only be re-sized afterwards
        e->rp[0].reserve(M * sizeof(u64), tot);
        e->rh[0].reserve(M, tot);
        e->rv[0].reserve(M * sizeof(u64), tot);
        u64* pos_all = e->rp[0].as<u64>();
        u8* head_all = e->rh[0].as<u8>();
        u64* suf_all = e->rv[0].as<u64>();
        if (m_r[me]) {
            PSAC_CUDA(cudaMemcpyAsync(pos_all + m_dsp[me], e->rp[1].p, m_r[me] * sizeof(u64), cudaMemcpyDeviceToDevice, st));
            PSAC_CUDA(cudaMemcpyAsync(head_all + m_dsp[me], e->rh[1].p, m_r[me], cudaMemcpyDeviceToDevice, st));
            PSAC_CUDA(cudaMemcpyAsync(suf_all + m_dsp[me], e->rv[1].p, m_r[me] * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        }
        PSAC_CUDA(cudaStreamSynchronize(st));  // the copies must have read the local lists before their buffers can move
        e->rp[1].reserve(M * sizeof(u64), tot);
        e->rh[1].reserve(M, tot);
        e->rv[1].reserve(M * sizeof(u64), tot);
        e->rk[0].reserve(M * sizeof(u64), tot);
        e->rk[1].reserve(M * sizeof(u64), tot);
        e->scratch.reserve(2 * M * sizeof(u64), tot);
        all_gather_v(e, C, pos_all, m_r, m_dsp, sizeof(u64));
        all_gather_v(e, C, head_all, m_r, m_dsp, 1);
        all_gather_v(e, C, suf_all, m_r, m_dsp, sizeof(u64));
        const int kb = (int)bits_for(n);
        u64 h = Cc;
        u64 m = M;
        int t = 1;  // output lists of the next round
        u64* ans = e->scratch.as<u64>();
        u64* vals0 = ans + M;
        e->lookback.reserve(std::max<size_t>(lookback_bytes(M), e->lookback.cap), tot);
        while (m > 0) {
            const int mbits = (int)bits_for(m - 1);
            if (kb + mbits > 64) throw arg_failure{"text too repetitive for a 64-bit round key"};
            const u64 ntiles = div_up(m, (size_t)RES_TILE);
            isa_answer_kernel<<<grid_for(e, m, 256, 8), 256, 0, st>>>(suf_all, m, h, n, ISA, text_lo, text_hi, ans);
            PSAC_NCCL(g_nccl.AllReduce(ans, ans, m, ncclUint64, ncclSum, C.comm, st));
            RoundKeyArgs K{};
            K.pos = pos_all;
            K.head = head_all;
            K.sa = nullptr;
            K.isa = nullptr;
            K.m = m;
            K.n = n;
            K.h = h;
            K.kbits = kb;
            K.keys = e->rk[0].as<u64>();
            K.vals = vals0;
            K.lb_max = e->lookback.as<u64>();
            K.tile_counter = e->counters() + 17;
            K.suf_in = suf_all;
            K.rank2 = ans;
            PSAC_CUDA(cudaMemsetAsync(K.lb_max, 0, ntiles * sizeof(u64), st));
            PSAC_CUDA(cudaMemsetAsync(K.tile_counter, 0, sizeof(u32), st));
            round_keys_kernel<u64><<<(unsigned)ntiles, RES_THREADS, 0, st>>>(K);
            e->launches += 2;
            PSAC_CUDA(cudaGetLastError());
            uint64_t sl2 = 0;
            // vals ping-pong: vals0 <-> suf_all is still needed by nothing after the keys are built, but keep it simple
            u64* vals1 = e->rv[t].as<u64>();
            const bool a2 = radix_sort_pairs<u64, u64>(e->radix_ws(), e->rk[0].as<u64>(), e->rk[1].as<u64>(), vals0, vals1, m, 0, kb + mbits, st, e->sm_count,
                                                       nullptr, &sl2);
            e->launches += sl2;
            // sorted suffixes must not alias the output list suf_out (= rv[t]): move them to vals0 if they ended in vals1
            const u64* sorted_keys = e->rk[a2 ? 1 : 0].as<u64>();
            if (a2) PSAC_CUDA(cudaMemcpyAsync(vals0, vals1, m * sizeof(u64), cudaMemcpyDeviceToDevice, st));
            ResolveArgs Q = R;
            Q.keys = sorted_keys;
            Q.vals = vals0;
            Q.pos_in = pos_all;
            Q.m = m;
            Q.isa = ISA;
            Q.bucket_out = nullptr;
            Q.halo = nullptr;
            Q.pos_base = 0;
            Q.pos_out = e->rp[t].p;
            Q.head_out = e->rh[t].as<u8>();
            Q.suf_out = e->rv[t].p;
            Q.cap = m;
            Q.lb_max = e->lookback.as<u64>();  // (the look-back buffer may have grown since R was filled in)
            Q.lb_sum = Q.lb_max + ntiles;
            Q.kbits = kb;
            Q.h = h;
            launch_resolve<u64, u64>(e, false, Q);
            u64 nb = 0;
            read_counts(e, &m, &nb);
            pos_all = e->rp[t].as<u64>();
            head_all = e->rh[t].as<u8>();
            suf_all = e->rv[t].as<u64>();
            t ^= 1;
            h *= 2;
            S.rounds += 1;
            if (h > 4 * n + 64 && m > 0) throw std::string("prefix doubling did not converge");
        }
        e->end(PH_ROUNDS);
    }

    // ---- S10 outputs: SA / LCP from key-range ownership [off, off + cnt) to exact blocks; ISA is already block-distributed
    e->begin(PH_OUTPUT);
    {
        std::vector<u64> scount(p), sdispl(p), rcount(p), rdispl(p);
        for (int b = 0; b < p; ++b) {
            const u64 lo = std::max(off, blk.start(b)), hi = std::min(off + cnt, blk.start(b) + blk.size(b));
            scount[b] = hi > lo ? hi - lo : 0;
            sdispl[b] = hi > lo ? lo - off : 0;
            const u64 rlo = std::max(off_key[b], text_lo), rhi = std::min(off_key[b + 1], text_hi);
            rcount[b] = rhi > rlo ? rhi - rlo : 0;
            rdispl[b] = rhi > rlo ? rlo - text_lo : 0;
        }
        const bool direct = index_bytes == 8;
        u64* blk_buf = direct ? nullptr : e->rk[0].as<u64>();
        if (!direct) e->rk[0].reserve((n_local + 16) * sizeof(u64), tot), blk_buf = e->rk[0].as<u64>();
        auto deliver = [&](const u64* src, void* dst) {
            u64* target = direct ? reinterpret_cast<u64*>(dst) : blk_buf;
            all_to_all_v(e, C, src, scount, sdispl, target, rcount, rdispl, sizeof(u64));
            if (!direct && n_local) {
                convert_kernel<u64, u32><<<grid_for(e, n_local, 256, 16), 256, 0, st>>>(target, reinterpret_cast<u32*>(dst), n_local);
                e->launches += 1;
            }
        };
        deliver(SA, sa_out);
        if (want_lcp) deliver(LCP, lcp_out);
        if (isa_out && n_local) {
            if (isa_inplace) {
                // already there
            } else if (direct)
                PSAC_CUDA(cudaMemcpyAsync(isa_out, ISA, n_local * sizeof(u64), cudaMemcpyDeviceToDevice, st));
            else {
                convert_kernel<u64, u32><<<grid_for(e, n_local, 256, 16), 256, 0, st>>>(ISA, reinterpret_cast<u32*>(isa_out), n_local);
                e->launches += 1;
            }
        }
        PSAC_CUDA(cudaGetLastError());
    }
    e->end(PH_OUTPUT);
    return true;
}

// ================================================================================================ sharded construction, v2
// Same interface and results as construct_sharded_core; a different schedule, built around ONE 64-bit word per suffix
// [carried key | global suffix index] and peer-visible memory:
//   1. every rank cuts the sort keys of ITS text block out of the replicated packed text and partitions them by the top
//      key digit; a digit's run is stored straight into the HBM of the rank that owns the digit (splitters = boundaries of
//      the all-gathered 256-bin digit histogram): the first digit pass of the sort IS the sample-sort exchange
//      (reference idxsort.hpp:22-83 -> samplesort.hpp:292-444), O(n/p) work per rank;
//   2. the remaining digits are sorted LSD inside the received top-digit segments, 8 bytes per suffix and pass;
//   3. heads / LCP of round 0 from the sorted words; SA -> ISA: every suffix's POSITION travels to the owner of its ISA
//      entry as one packed word (peer stores fused into the owner partition), the few unresolved suffixes are fixed up to
//      their bucket head afterwards;
//   4. later rounds are DISTRIBUTED: a bucket never straddles two ranks, so every rank sorts its own unresolved suffixes;
//      ISA[s + h] is read from, and new bucket ids are written to, the owners' ISA blocks through peer memory
//      (reference bulk_rma.hpp:112-135 / suffix_array.hpp:1032-1285 without the all-to-all);
//   5. SA is pulled from the key-range owners into exact blocks by a copy kernel over peer memory, LCP moves with
//      grouped ncclSend / ncclRecv, ISA is already block-distributed.
// Returns false -- identically on all ranks -- when it cannot take the input (no peer access, sigma = 256 quirk, top digits
// too skewed to balance, exchange word does not fit 64 bits); the caller then runs construct_sharded_core.

// Host plan of the fused first pass (no GPU needed; exposed for the CPU tests as psacb200_plan_word_exchange).
// cnt[s * nb + d] = suffixes of text block s whose top key digit is d.  Owner o sorts the digits [first[o], first[o+1]).
// Inside a segment the runs of the sources are laid out in the order p-1, 0, 1, .., p-2: the suffixes that run past the end
// of the text belong to the last block and must precede equal keys (stable passes keep them there).
struct WordExchangePlan {
    int p = 0, nb = 0;
    std::vector<size_t> first;              // [p + 1]
    std::vector<u64> cnt_key, off_key;      // [p], [p + 1]
    std::vector<int> owner;                 // [nb]
    std::vector<u64> seg_dense, seg_pad;    // [p][257] (entries >= nb repeat the total)
    std::vector<u64> run_off;               // [p][nb]: element offset of source s's run of digit d inside the OWNER's padded layout
    u64 max_cnt = 0, max_pad = 0;
    bool balanced = false;
};

static void plan_word_exchange(const u64* cnt, int p, int nb, u64 n, u64 pad_tile, WordExchangePlan& P) {
    P.p = p;
    P.nb = nb;
    std::vector<u64> tot(nb, 0);
    for (int s = 0; s < p; ++s)
        for (int d = 0; d < nb; ++d) tot[d] += cnt[(size_t)s * nb + d];
    choose_splitters(tot.data(), nb, n, p, P.first, P.cnt_key);
    P.off_key.assign(p + 1, 0);
    P.owner.assign(nb, 0);
    P.seg_dense.assign((size_t)p * 257, 0);
    P.seg_pad.assign((size_t)p * 257, 0);
    P.run_off.assign((size_t)p * nb, 0);
    P.max_cnt = P.max_pad = 0;
    P.balanced = true;
    const u64 cap = 2 * ((n + p - 1) / (u64)p) + 4096;
    for (int o = 0; o < p; ++o) {
        P.off_key[o + 1] = P.off_key[o] + P.cnt_key[o];
        if (P.cnt_key[o] == 0 || P.cnt_key[o] > cap) P.balanced = false;
        P.max_cnt = std::max(P.max_cnt, P.cnt_key[o]);
        u64 dense = 0, pad = 0;
        for (int d = 0; d <= 256; ++d) {
            P.seg_dense[(size_t)o * 257 + d] = dense;
            P.seg_pad[(size_t)o * 257 + d] = pad;
            if (d < nb && (size_t)d >= P.first[o] && (size_t)d < P.first[o + 1]) {
                P.owner[d] = o;
                u64 run = pad;
                for (int i = 0; i < p; ++i) {
                    const int src = (i + p - 1) % p;  // p-1, 0, 1, .., p-2
                    P.run_off[(size_t)src * nb + d] = run;
                    run += cnt[(size_t)src * nb + d];
                }
                dense += tot[d];
                pad += (tot[d] + pad_tile - 1) / pad_tile * pad_tile;
            }
        }
        P.max_pad = std::max(P.max_pad, pad);
    }
}

// ------------------------------------------------------------------------------------------------ v2, digit pass 1
// The packed text is replicated, so a rank needs no exchange to collect the suffixes it sorts: it scans the whole text and
// keeps the suffixes whose top key digit lies in its digit range [dlo, dhi) -- the selection IS the first (top-digit) pass
// of the sort.  One kernel, no histogram pre-pass: a CTA takes 512 x 64 consecutive characters, marks the selected ones in a
// 64-bit mask per thread, and in batches of <= SELW_CAP elements ranks them by digit in shared memory (order inside a
// segment is free: equal keys are one bucket anyway), reserves every digit's run with ONE atomic on the segment's cursor
// and writes the runs out coalesced as words [carried key | suffix index] into the tile-padded segment layout.  The
// suffixes that run past the end of the text are placed first in their segments by select_tail_words_kernel (they must
// precede equal keys, shortest first, and the LSD passes that follow are stable).
// Cost per rank: reads n * lbits / 8 bytes of text, ~6 integer instructions per character, writes 8 bytes per selected
// suffix; no NVLink traffic (measured alternative: the same pass with peer stores into the owners' HBM, 13.3 ms for
// 2^30 suffixes on 2 GPUs against 716 GB/s of achievable NVLink store bandwidth, profiles/r2_peer_bw.txt).
constexpr int SELW_THREADS = 512;
constexpr int SELW_CPT = 64;      // characters per thread (32 when few ranks share the text: a CTA's selection then fits one batch)
constexpr int SELW_CAP = 8192;    // staged words per batch
constexpr size_t SELW_SMEM = SELW_CAP * sizeof(u64) + (SELW_THREADS / 32) * RADIX * sizeof(u32);

struct SelectWordsArgs {
    const u64* stream;
    u64 n_main;                   // suffixes 0 .. n_main-1 (the others run past the end: select_tail_words_kernel)
    int K, tb, cb, ib;            // key bits, top digit bits, carried bits, bits of the index field
    WordIdx widx;                 // the index field: (owner, block-local index)
    u32 dlo, dhi;
    const u64* seg_pad;           // [257] padded segment starts of this rank
    unsigned long long* cursor;   // [256] elements placed so far in every segment
    u64* out;
};

template <int WPT, int LBITS>
__device__ __forceinline__ u64 selw_bits(const u64 (&wd)[WPT + 1], int c) {
    const int o = c * LBITS, j = o >> 6, off = o & 63;
    u64 hi = wd[0], lo = wd[1];
#pragma unroll
    for (int q = 1; q < WPT; ++q)
        if (q == j) {
            hi = wd[q];
            lo = wd[q + 1];
        }
    return off ? ((hi << off) | (lo >> (64 - off))) : hi;
}

template <int LBITS, int CPT>
__global__ void __launch_bounds__(SELW_THREADS, 2) select_words_kernel(SelectWordsArgs A) {
    static_assert(CPT == 64 || (CPT == 32 && LBITS >= 2), "a thread takes whole stream words");
    extern __shared__ __align__(16) u64 sw_stage[];  // SELW_CAP words, then the per-warp digit tables
    constexpr int NW = SELW_THREADS / 32;
    u32* tab = reinterpret_cast<u32*>(sw_stage + SELW_CAP);  // [NW][RADIX]: counts, then running positions
    __shared__ u8 s_dig[SELW_CAP];
    __shared__ u64 s_goff[RADIX];
    __shared__ u32 s_wsum[NW];
    constexpr int WPT = LBITS * CPT / 64;  // CPT characters = WPT words
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32* mytab = tab + warp * RADIX;
    const u64 g0 = ((u64)blockIdx.x * SELW_THREADS + tid) * CPT;
    u64 wd[WPT + 1];
#pragma unroll
    for (int j = 0; j <= WPT; ++j) wd[j] = 0;
    for (int e = tid; e < NW * RADIX; e += SELW_THREADS) tab[e] = 0u;
    __syncthreads();
    u64 mask = 0;
    if (g0 < A.n_main) {
        const u64 w0 = (g0 * LBITS) >> 6;
#pragma unroll
        for (int j = 0; j <= WPT; ++j) wd[j] = __ldg(A.stream + w0 + j);  // (the stream carries two zero words of padding)
        const u32 span = A.dhi - A.dlo;
        const u64 left = A.n_main - g0;
        const u64 vmask = left >= CPT ? (CPT == 64 ? ~0ull : ((1ull << (CPT & 63)) - 1ull)) : ((1ull << left) - 1ull);  // characters of this thread that are suffixes < n_main
        const int dsh = 32 - A.tb;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int o = c * LBITS, j = o >> 6, off = o & 63;  // compile-time after unrolling
            // the top 32 bits of the window at this character are enough for the digit
            const u32 a = off < 32 ? (u32)(wd[j] >> 32) : (u32)wd[j];
            const u32 b = off < 32 ? (u32)wd[j] : (u32)(wd[j + 1] >> 32);
            const u32 v32 = __funnelshift_l(b, a, off & 31);
            const u32 d = v32 >> dsh;
            if (d - A.dlo < span && ((vmask >> c) & 1ull)) {
                mask |= 1ull << c;
                atomicAdd(&mytab[d], 1u);  // digit counts of the whole CTA (valid when everything fits one batch)
            }
        }
    }
    // sequence numbers of the selected characters in thread order
    const u32 mine = (u32)__popcll(mask);
    const u32 incl = warp_inclusive_sum_u32(mine);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u32 tseq = incl - mine, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w < warp) tseq += s_wsum[w];
        total += s_wsum[w];
    }
    const bool single = total <= SELW_CAP;
    const u64 keymask = A.cb >= 64 ? ~0ull : ((1ull << A.cb) - 1ull);
    // index field of my first character; my characters span at most two text blocks (blocks hold >= 2^16 characters)
    u64 loc0 = 0, bsize = ~0ull;
    u32 own0 = 0;
    if (mask) {
        own0 = A.widx.div.owner(g0, &loc0);
        bsize = A.widx.block_size(own0);
    }
    for (u32 lo = 0; lo < total; lo += SELW_CAP) {
        const u32 hi = lo + SELW_CAP < total ? lo + SELW_CAP : total;
        if (!single) {
            // more than one batch (few ranks: a large share of the characters is mine): recount per batch
            __syncthreads();
            for (int e = tid; e < NW * RADIX; e += SELW_THREADS) tab[e] = 0u;
            __syncthreads();
            u64 mm = mask;
            u32 seq = tseq;
            while (mm) {
                const int c = __ffsll((long long)mm) - 1;
                mm &= mm - 1;
                if (seq >= lo && seq < hi) atomicAdd(&mytab[(u32)(selw_bits<WPT, LBITS>(wd, c) >> (64 - A.tb))], 1u);
                ++seq;
            }
        }
        __syncthreads();
        // per digit: exclusive offsets of the warps, exclusive scan over the digits; ONE global atomic per non-empty digit
        // reserves its run in the segment
        u32 cnt = 0, inc = 0;
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const u32 cw = tab[w * RADIX + tid];
                tab[w * RADIX + tid] = cnt;
                cnt += cw;
            }
            inc = warp_inclusive_sum_u32(cnt);
            if (lane == 31) s_wsum[warp] = inc;
        }
        __syncthreads();
        if (tid < RADIX) {
            u32 pre = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; ++w) pre += (w < warp) ? s_wsum[w] : 0u;
            const u32 bstart = pre + inc - cnt;
#pragma unroll
            for (int w = 0; w < NW; ++w) tab[w * RADIX + tid] += bstart;
            if (cnt) s_goff[tid] = A.seg_pad[tid] + (u64)atomicAdd(&A.cursor[tid], (unsigned long long)cnt) - (u64)bstart;
        }
        __syncthreads();
        {  // place: any order inside a digit
            u64 mm = mask;
            u32 seq = tseq;
            while (mm) {
                const int c = __ffsll((long long)mm) - 1;
                mm &= mm - 1;
                if (seq >= lo && seq < hi) {
                    const u64 v = selw_bits<WPT, LBITS>(wd, c);
                    const u32 d = (u32)(v >> (64 - A.tb));
                    const u32 p = atomicAdd(&mytab[d], 1u);
                    const u64 loc = loc0 + (u64)c;
                    const u64 fld = loc < bsize ? (((u64)own0 << A.widx.lb) | loc) : (((u64)(own0 + 1) << A.widx.lb) | (loc - bsize));
                    sw_stage[p] = (((v >> (64 - A.K)) & keymask) << A.ib) | fld;
                    s_dig[p] = (u8)d;
                }
                ++seq;
            }
        }
        __syncthreads();
        for (u32 s = tid; s < hi - lo; s += SELW_THREADS) st_stream(A.out + s_goff[s_dig[s]] + s, sw_stage[s]);
        __syncthreads();
    }
}

// the T <= 63 suffixes that run past the end of the text: first in their segments, shortest first; also initialises the cursors
__global__ void __launch_bounds__(64) select_tail_words_kernel(SelectWordsArgs A, u64 n, u64 T, int lbits) {
    __shared__ u32 s_d[64];
    const int t = threadIdx.x;
    u32 d = ~0u;
    u64 key = 0;
    const u64 g = n - 1 - (u64)t;
    if ((u64)t < T) {
        key = stream_extract(A.stream, g, lbits, A.K);
        const u32 dd = (u32)(key >> A.cb);
        if (dd - A.dlo < A.dhi - A.dlo) d = dd;
    }
    s_d[t] = d;
    __syncthreads();
    if (d != ~0u) {
        u32 before = 0, all = 0;
        for (int i = 0; i < 64; ++i) {
            if (s_d[i] == d) {
                all += 1;
                if (i < t) before += 1;
            }
        }
        const u64 keymask = A.cb >= 64 ? ~0ull : ((1ull << A.cb) - 1ull);
        A.out[A.seg_pad[d] + before] = ((key & keymask) << A.ib) | A.widx.encode(g);
        if (before == 0) A.cursor[d] = all;  // (the cursors were zeroed before)
    }
}

// histogram of the top key digit of the suffixes [g_lo, g_lo + cnt) (one rank's text block)
template <int LBITS>
__global__ void __launch_bounds__(512) digit_hist_kernel(const u64* __restrict__ stream, u64 g_lo, u64 cnt, int tb, u64* __restrict__ hist) {
    __shared__ u32 sh[4][RADIX];
    for (int e = threadIdx.x; e < 4 * RADIX; e += blockDim.x) (&sh[0][0])[e] = 0;
    __syncthreads();
    u32* my = sh[(threadIdx.x >> 5) & 3];
    constexpr int WPT = LBITS;
    for (u64 t0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * SELW_CPT; t0 < cnt; t0 += (u64)gridDim.x * blockDim.x * SELW_CPT) {
        const u64 g0 = g_lo + t0;
        const u64 bit0 = g0 * LBITS, w0 = bit0 >> 6;
        const int sh0 = (int)(bit0 & 63);  // a block need not start on a word boundary
        u64 wd[WPT + 2];
#pragma unroll
        for (int j = 0; j <= WPT + 1; ++j) wd[j] = __ldg(stream + w0 + j);
        if (sh0) {
#pragma unroll
            for (int j = 0; j <= WPT; ++j) wd[j] = (wd[j] << sh0) | (wd[j + 1] >> (64 - sh0));
        }
#pragma unroll
        for (int c = 0; c < SELW_CPT; ++c) {
            const int o = c * LBITS, j = o >> 6, off = o & 63;
            const u64 v = off ? ((wd[j] << off) | (wd[j + 1] >> (64 - off))) : wd[j];
            if (t0 + c < cnt) atomicAdd(&my[(u32)(v >> (64 - tb))], 1u);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < RADIX; c += blockDim.x) {
        const u64 v = (u64)sh[0][c] + sh[1][c] + sh[2][c] + sh[3][c];
        if (v) atomicAdd((unsigned long long*)&hist[c], (unsigned long long)v);
    }
}

template <typename F>
void dispatch_lbits(int lbits, F&& f) {
    switch (lbits) {
        case 1: f(std::integral_constant<int, 1>()); break;
        case 2: f(std::integral_constant<int, 2>()); break;
        case 4: f(std::integral_constant<int, 4>()); break;
        default: f(std::integral_constant<int, 8>()); break;
    }
}

// key source of the SA -> ISA step of v2: element g of the sorted words is suffix (word & mask) at SA position off + g; what
// travels is [index inside the owner's ISA block | rank field | g] (see OwnerPackSrcT).  The pass partitions by
// digit = (owning rank, middle bits of the block-local index): one run per (destination, ISA window group), so that the
// receiver's single partition pass by the top index bits ends with windows of 2^-(8 + bshift) of the block -- small enough
// to stay in L2 while the final scatter fills them.
struct OwnerMidSrc {
    using Stage = u64;
    using Out = u64;
    static constexpr bool FROM_TEXT = false;
    static constexpr bool PEER = false;
    const u64* __restrict__ win;
    u64 fmask, lmask;    // masks of the index field of a sort word and of its block-local part
    int lb;
    int rank_shift, idx_shift;
    u32 rank_mask;
    u64 me;
    int mshift, bshift;  // digit = owner << bshift | (local >> mshift) & ((1 << bshift) - 1)
    __device__ __forceinline__ Stage load_key(size_t g) const {
        const u64 f = ld_stream(win + g) & fmask;
        return ((f & lmask) << idx_shift) | ((f >> lb) << rank_shift) | (u64)g;
    }
    __device__ __forceinline__ u32 digit(Stage k) const {
        const u32 owner = (u32)(k >> rank_shift) & rank_mask;
        return (owner << bshift) | ((u32)((k >> idx_shift) >> mshift) & ((1u << bshift) - 1u));
    }
    __device__ __forceinline__ u32 hist_digit(size_t g) const {
        const u64 f = ld_stream(win + g) & fmask;
        return ((u32)(f >> lb) << bshift) | ((u32)((f & lmask) >> mshift) & ((1u << bshift) - 1u));
    }
    __device__ __forceinline__ Out out_key(Stage k) const { return (k & ~((u64)rank_mask << rank_shift)) | (me << rank_shift); }
    __device__ __forceinline__ NoVal load_val(size_t) const { return NoVal(); }
    __device__ __forceinline__ u8 load_aux(size_t, Stage) const { return 0; }
};

// complete key and suffix of the last sorted word of a rank (the halo of the next rank's first boundary)
__global__ void __launch_bounds__(32) last_word_kernel(const u64* __restrict__ words, u64 cnt, u64 top_digit, int cb, int ib, WordIdx widx, u64* __restrict__ out) {
    if (threadIdx.x == 0) {
        const u64 w = cnt ? words[cnt - 1] : 0;
        out[0] = (top_digit << cb) | (w >> ib);
        out[1] = widx.decode(w);
    }
}

// SA block [dst_lo, dst_lo + m) of this rank, pulled from the sorted words of the key-range owners (peer memory):
// global SA position g lives at word (g - off_key[a]) of owner a.
struct PullArgs {
    const u64* src[16];
    u64 off_key[17];
    int p;
    u64 dst_lo, m;
    WordIdx widx;
    void* dst;
};
template <typename OutT>
__global__ void __launch_bounds__(256) pull_sa_kernel(PullArgs A) {
    OutT* dst = reinterpret_cast<OutT*>(A.dst);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < A.m; i += (u64)gridDim.x * blockDim.x) {
        const u64 g = A.dst_lo + i;
        int a = 0;
        while (a + 1 < A.p && A.off_key[a + 1] <= g) ++a;
        st_stream(dst + i, (OutT)A.widx.decode(A.src[a][g - A.off_key[a]]));
    }
}

bool construct_sharded_v2(psacb200_engine* e, const ShardComm& C, const u8* d_text_local, u64 n_local, u64 n, int index_bytes, unsigned flags, unsigned k,
                          void* sa_out, void* isa_out, void* lcp_out) {
    const bool want_lcp = (flags & PSACB200_LCP) != 0;
    const int p = C.world, me = C.rank;
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    psacb200_stats& S = e->stats;
    const BlkDist blk(n, p);
    if (blk.size(me) != n_local) throw arg_failure{"the input text must be equally block decomposed across all ranks (reference suffix_array.hpp:226)"};
    if (e->peer_map == nullptr || getenv("PSACB200_SHARDED_V1")) return false;
    S.internal_index_bytes = 8;
    S.sort_elt_bytes = 8;
    e->v1_stats = false;

    // ---- alphabet, replicated packed text
    Alphabet alpha;
    prepare_text_sharded(e, C, d_text_local, n_local, n, alpha);
    if (alpha.zero_code_used) return false;  // (the sigma = 256 quirk takes the generic resolve of construct_sharded_core)
    const int lbits = alpha.lbits;
    u64* stream = e->packed.as<u64>();

    e->mark("text");
    // ---- key shape: K = tb (top digit) + cb (carried) bits, cb + ib <= 64
    // index field of a sort word: (owner of the suffix's ISA entry, index inside that rank's block)
    const int idx_bits = std::max(1, (int)bits_for(blk.size(0) ? blk.size(0) - 1 : 0)), rank_bits = std::max(1, (int)bits_for((u64)p - 1));
    const int ib = idx_bits + rank_bits;
    WordIdx widx;
    widx.div = BlkDiv::make(n, p);
    widx.lb = idx_bits;
    widx.mask = (1ull << ib) - 1ull;
    // The key is the first K BITS of the suffix in the packed text -- not necessarily whole characters: with 2-bit DNA and a
    // 33-bit index field 31 carried + 8 top bits = 39 bits = 19.5 characters, and the half character halves the number of
    // suffixes that stay unresolved.  Cc = characters the key covers completely (what a shared bucket is known to agree on),
    // Ct = characters it touches (suffixes shorter than that run past the end of the text inside the key).
    const unsigned Cmax = choose_key_chars(n, lbits, k);
    int K = (int)Cmax * lbits;
    {
        const int tb0 = std::min(RADIX_BITS, K);
        K = std::min(K, tb0 + (64 - ib));
        if (const char* force = getenv("PSACB200_V2_KEYBITS")) {  // test knob: a key that ends inside a character at small sizes
            const int f = atoi(force);
            if (f >= lbits && f <= K) K = f;
        }
    }
    const int tb = std::min(RADIX_BITS, K), cb = K - tb;
    if (K < lbits || cb > 64 - ib) return false;  // (less than one character of key: not worth a sharded sort)
    const unsigned Cc = (unsigned)(K / lbits), Ct = (unsigned)((K + lbits - 1) / lbits);
    const int nb = 1 << tb;
    const u64 T = (n < (u64)Ct - 1) ? n : (u64)Ct - 1;
    if (T > blk.size(p - 1)) return false;  // the suffixes that run past the end must all lie in the last block
    S.key_chars = Cc;

    // ---- digit histogram of my text block, all-gathered: segment sizes and digit ownership (host plan)
    const size_t LTILE = word_sort_tile();
    e->begin(PH_SORT);
    e->segws.reserve((2 * (RADIX + 1) + RADIX * RADIX) * sizeof(u64) + 256 + (size_t)(p + 2) * RADIX * sizeof(u64), tot);
    u64* d_segtab = e->segws.as<u64>();                              // seg_dense[257], seg_pad[257], segbase[256][256]
    u64* d_cnt = d_segtab + 2 * (RADIX + 1) + RADIX * RADIX + 32;    // [p][256] digit counts of every block
    unsigned long long* d_cursor = reinterpret_cast<unsigned long long*>(d_cnt + (size_t)p * RADIX);  // [256]
    PSAC_CUDA(cudaMemsetAsync(d_cnt + (size_t)me * RADIX, 0, RADIX * sizeof(u64), st));
    PSAC_CUDA(cudaMemsetAsync(d_cursor, 0, RADIX * sizeof(u64), st));
    dispatch_lbits(lbits, [&](auto LB) {
        digit_hist_kernel<decltype(LB)::value><<<grid_for(e, n_local / SELW_CPT + 1, 512, 4), 512, 0, st>>>(stream, blk.start(me), n_local, tb, d_cnt + (size_t)me * RADIX);
    });
    e->launches += 1;
    PSAC_CUDA(cudaGetLastError());
    e->mark("p1_hist");
    PSAC_NCCL(g_nccl.AllGather(d_cnt + (size_t)me * RADIX, d_cnt, RADIX, ncclUint64, C.comm, st));
    std::vector<u64> h_cnt((size_t)p * RADIX);
    PSAC_CUDA(cudaMemcpyAsync(h_cnt.data(), d_cnt, h_cnt.size() * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    std::vector<u64> cnt_sd((size_t)p * nb);
    for (int s = 0; s < p; ++s)
        for (int d = 0; d < nb; ++d) cnt_sd[(size_t)s * nb + d] = h_cnt[(size_t)s * RADIX + d];
    WordExchangePlan P;
    plan_word_exchange(cnt_sd.data(), p, nb, n, cb > 0 ? (u64)LTILE : 1, P);
    if (!P.balanced) return false;  // top key digits too skewed for digit-boundary splitters
    if (P.max_cnt >= (1ull << 32)) return false;  // (the heads kernel keeps positions inside a rank's range in 32 bits)
    const u64 cnt = P.cnt_key[me], off = P.off_key[me];
    // exchange word of the SA -> ISA step: [local index | rank | position relative to the sender]
    const int rel_bits = std::max(1, (int)bits_for(P.max_cnt - 1));
    if (rel_bits + rank_bits + idx_bits > 64) return false;

    // ---- peer-visible buffers (identical offsets on all ranks)
    const size_t PADE = align_up((size_t)P.max_pad + LTILE, LTILE) + LTILE;  // padded word buffers
    const size_t wbytes = align_up(PADE * sizeof(u64), 256), rbytes = align_up((blk.size(0) + 16) * sizeof(u64), 256);
    if (!arena_ensure(e, C, 2 * wbytes + 3 * rbytes)) return false;
    PeerArena& A = *reinterpret_cast<PeerArena*>(e->peer_map);
    auto at = [&](int r, size_t o) { return reinterpret_cast<u64*>(A.peer[r] + o); };
    const size_t oW[2] = {0, wbytes}, oR0 = 2 * wbytes, oR1 = 2 * wbytes + rbytes, oISA = 2 * wbytes + 2 * rbytes;
    u64* W[2] = {at(me, oW[0]), at(me, oW[1])};
    u64* ISA = at(me, oISA);
    S.reserved = 1u;
    e->lookback.reserve(lookback_bytes(std::max<u64>(PADE, n_local) + 4 * LTILE), tot);
    RadixWorkspace ws = e->radix_ws();
    const size_t rows = (size_t)(P.seg_pad[(size_t)me * 257 + 256] / LTILE) + 1;
    SegWorkspace sw;
    sw.seg_dense = d_segtab;
    sw.seg_pad = sw.seg_dense + (RADIX + 1);
    sw.segbase = sw.seg_pad + (RADIX + 1);
    e->tb[0].reserve(rows * sizeof(u32) + 256, tot);
    sw.tile_info = e->tb[0].as<u32>();
    // my segment tables
    u64* hp = e->h_pinned + 512;  // [0..256] seg_dense, [257..513] seg_pad
    for (int d = 0; d <= 256; ++d) {
        hp[d] = P.seg_dense[(size_t)me * 257 + d];
        hp[257 + d] = P.seg_pad[(size_t)me * 257 + d];
    }
    PSAC_CUDA(cudaMemcpyAsync(sw.seg_dense, hp, 2 * 257 * sizeof(u64), cudaMemcpyHostToDevice, st));
    e->mark("p1_plan");

    // ---- pass 1: select my digit range out of the whole (replicated) text, straight into the padded segment layout
    rank_barrier(e, C);  // the peers have pulled their SA blocks out of my word buffers (previous call)
    e->mark("p1_barrier");
    {
        SelectWordsArgs SA_{};
        SA_.stream = stream;
        SA_.n_main = n - T;
        SA_.K = K;
        SA_.tb = tb;
        SA_.cb = cb;
        SA_.ib = ib;
        SA_.widx = widx;
        SA_.dlo = (u32)P.first[me];
        SA_.dhi = (u32)P.first[me + 1];
        SA_.seg_pad = sw.seg_pad;
        SA_.cursor = d_cursor;
        SA_.out = W[0];
        select_tail_words_kernel<<<1, 64, 0, st>>>(SA_, n, T, lbits);
        // 64 characters per thread; 32 where a rank keeps a large share of the text (fewer than 4 ranks), so that the
        // selection of a CTA still fits one batch of the staging buffer
        const bool half = p < 4 && lbits >= 2;
        const u64 ctas = div_up(SA_.n_main ? SA_.n_main : 1, (size_t)SELW_THREADS * (half ? 32 : 64));
        dispatch_lbits(lbits, [&](auto LB) {
            constexpr int L = decltype(LB)::value;
            static bool seen[64] = {};
            const bool first = first_use_on_device(seen);
            auto launch = [&](auto kern) {
                if (first) PSAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_SMEM));
                kern<<<(unsigned)ctas, SELW_THREADS, SELW_SMEM, st>>>(SA_);
            };
            if constexpr (L >= 2) {
                if (first) PSAC_CUDA(cudaFuncSetAttribute(select_words_kernel<L, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_SMEM));
                if (half) {
                    select_words_kernel<L, 32><<<(unsigned)ctas, SELW_THREADS, SELW_SMEM, st>>>(SA_);
                    return;
                }
            }
            launch(select_words_kernel<L, 64>);
        });
        e->launches += 2;
        PSAC_CUDA(cudaGetLastError());
    }
    cudaEventRecord(e->ev_end[PH_PASS1], st);
    e->mark("p1_select");

    // ---- the remaining digits, LSD inside my segments
    int x = 0;
    RadixPlan plan_used{};
    if (cb > 0) {
        uint64_t sl = 0;
        x = radix_sort_words_seg(ws, sw, rows, W, ib, cb, st, &plan_used, &sl, e->ev_scatter);
        e->launches += sl;
        e->scatter_passes = plan_used.npass;
    }
    S.sort_passes = 1 + plan_used.npass;
    e->end(PH_SORT);
    e->mark("lsd");
    u64* Wx = W[x];  // sorted words, dense: SA positions [off, off + cnt)

    // ---- SA -> ISA, part A (before the heads: it needs the sorted words only -- every suffix sends its POSITION, the
    //      unresolved ones are fixed up to their bucket head afterwards).  Local partition of the exchange words by
    //      (owner, window group) into the spare word buffer; the bin starts of all ranks are all-gathered.
    e->begin(PH_ISA);
    const int rank_shift = rel_bits, idx_shift = rel_bits + rank_bits;
    const int bshift = RADIX_BITS - rank_bits;                       // window groups per owner = 1 << bshift
    const int hshift = idx_bits > RADIX_BITS ? idx_bits - RADIX_BITS : 0;  // receiver's pass: top 8 bits of the local index
    const int mshift = hshift > bshift ? hshift - bshift : 0;
    u64* Sbuf = W[1 - x];
    u64* h_bstart = e->h_pinned + 2048;  // [p][256] bin starts of every rank's partition (pinned)
    {
        OwnerMidSrc src{Wx, widx.mask, (1ull << idx_bits) - 1ull, idx_bits, rank_shift, idx_shift, (1u << rank_bits) - 1u, (u64)me, mshift, bshift};
        launch_pass<OwnerMidSrc, NoVal, false>(ws, src, Sbuf, nullptr, nullptr, cnt, st);
        e->launches += LAUNCHES_PER_PASS;
        PSAC_CUDA(cudaGetLastError());
        e->mark("isa_partA");
        PSAC_CUDA(cudaMemcpyAsync(d_cnt + (size_t)me * RADIX, ws.gbase, RADIX * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        PSAC_NCCL(g_nccl.AllGather(d_cnt + (size_t)me * RADIX, d_cnt, RADIX, ncclUint64, C.comm, st));
        PSAC_CUDA(cudaMemcpyAsync(h_bstart, d_cnt, (size_t)p * RADIX * sizeof(u64), cudaMemcpyDeviceToHost, st));
        PSAC_CUDA(cudaEventRecord(e->ev_x[0], st));
    }
    e->mark("isa_owner");

    // ---- resolve round 0 (heads, LCP, unresolved list) on the sorted words
    e->begin(PH_RESOLVE);
    u64* d_last = e->shard_meta();  // [2 * p] all-gathered {last complete key, last suffix}
    int last_digit = 0;
    for (int d = 0; d < nb; ++d)
        if (P.owner[d] == me && P.seg_dense[(size_t)me * 257 + d + 1] > P.seg_dense[(size_t)me * 257 + d]) last_digit = d;
    last_word_kernel<<<1, 32, 0, st>>>(Wx, cnt, (u64)last_digit, cb, ib, widx, d_last + 2 * me);
    PSAC_NCCL(g_nccl.AllGather(d_last + 2 * me, d_last, 2, ncclUint64, C.comm, st));
    if (want_lcp) e->lcp.reserve((cnt + 16) * sizeof(u64), tot);
    u64* LCP = want_lcp ? e->lcp.as<u64>() : nullptr;
    // 64-bit caller: the LCP of the positions of my key range that lie inside my own output block is written in place;
    // only the slivers that belong to the neighbours' blocks go through the key-range buffer and the re-balancing step
    const u64 blk_lo = blk.start(me), blk_hi = blk_lo + n_local;
    const bool lcp_direct = want_lcp && index_bytes == 8 && std::min(off + cnt, blk_hi) > std::max(off, blk_lo);
    const u64 main_lo = lcp_direct ? std::max(off, blk_lo) - off : 0, main_hi = lcp_direct ? std::min(off + cnt, blk_hi) - off : 0;
    u64* lcp_main = lcp_direct ? reinterpret_cast<u64*>(lcp_out) + ((long long)off - (long long)blk_lo) : nullptr;
    u64 ucap = std::max<u64>(unresolved_cap(cnt), 1);
    auto reserve_lists = [&](int t, u64 cap_) {
        e->rp[t].reserve(cap_ * sizeof(u64), tot);
        e->rh[t].reserve(cap_, tot);
        e->rv[t].reserve(cap_ * sizeof(u64), tot);
    };
    reserve_lists(1, ucap);
    TailList* tails = reinterpret_cast<TailList*>(e->tail_list());
    tail_positions_kernel<u64><<<1, 64, 0, st>>>(Wx, cnt, sw.seg_dense, cb, stream, n, T, lbits, K, tb, (u32)P.first[me], (u32)P.first[me + 1], tails, ib);
    const u64 htiles = div_up(cnt, (size_t)HD_TILE);
    HeadsArgs H{};
    H.keys = Wx;
    H.seg_dense = sw.seg_dense;
    H.seg_shift = cb;
    {
        u8* sflags = e->lookback.as<u8>() + 2 * htiles * sizeof(u64);  // behind the tile aggregates
        PSAC_CUDA(cudaMemsetAsync(sflags, 0, htiles, st));
        seg_flag_kernel<<<1, 256, 0, st>>>(sw.seg_dense, cnt, (u64)HD_TILE, sflags);
        if (me > 0) PSAC_CUDA(cudaMemsetAsync(sflags, 1, 1, st));  // position 0 is compared with the previous rank's last complete key
        H.seg_flags = sflags;
    }
    H.vals = nullptr;
    H.m = cnt;
    H.n = n;
    H.lbits = lbits;
    H.C = (int)Cc;
    H.kbits = K;
    H.tails = tails;
    H.bucket_out = nullptr;
    H.isa = nullptr;
    H.lcp = LCP;
    H.lcp_main = lcp_main;
    H.main_lo = main_lo;
    H.main_hi = main_hi;
    H.pos_out = e->rp[1].p;
    H.head_out = e->rh[1].as<u8>();
    H.suf_out = e->rv[1].p;
    H.cap = ucap;
    H.counts = e->counts();
    H.pos_base = 0;  // positions are relative to my first SA position `off`
    H.halo = me > 0 ? d_last + 2 * (me - 1) : nullptr;
    H.agg_max = e->lookback.as<u64>();
    H.agg_sum = H.agg_max + htiles;
    H.word_shift = ib;
    H.widx = widx;
    PSAC_CUDA(cudaMemsetAsync(H.counts, 0, 2 * sizeof(u64), st));
    heads_kernel<u64, u64, 0, true><<<(unsigned)htiles, HD_THREADS, 0, st>>>(H);
    tile_scan_kernel<<<1, 1024, 0, st>>>(H.agg_max, H.agg_sum, htiles, H.counts);
    heads_kernel<u64, u64, 1, true><<<(unsigned)htiles, HD_THREADS, 0, st>>>(H);
    e->launches += 6;
    PSAC_CUDA(cudaGetLastError());
    e->end(PH_RESOLVE);
    e->mark("heads");
    // ---- SA -> ISA, part X (queued now, while the heads kernels run): every (destination, window group) run goes to the
    //      destination's receive buffer with the copy engines (peer-mapped memory), group-major / source-minor
    {
        cudaStream_t cs = e->copy_streams[0];
        PSAC_CUDA(cudaEventSynchronize(e->ev_x[0]));  // (the heads kernels are queued behind it and keep the GPU busy)
        const int B = 1 << bshift;
        auto count_of = [&](int s_, int bin) {
            const u64 a0 = h_bstart[(size_t)s_ * RADIX + bin];
            const u64 a1 = bin + 1 < RADIX ? h_bstart[(size_t)s_ * RADIX + bin + 1] : P.cnt_key[s_];
            return a1 - a0;
        };
        // several copy streams (= copy engines) side by side: the local share on one, the peers spread over the others,
        // every rank starting with its right neighbour so that no destination is hit by all senders at once
        constexpr int NCS = psacb200_engine::COPY_STREAMS;
        for (int i = 0; i < NCS; ++i) PSAC_CUDA(cudaStreamWaitEvent(e->copy_streams[i], e->ev_x[0], 0));
        PSAC_CUDA(cudaEventRecord(e->ev_xt[0], cs));
        std::vector<u64> dst_off((size_t)p * B, 0);  // where my run of (destination b, group) starts in b's receive buffer
        for (int b = 0; b < p; ++b) {
            u64 run = 0;
            for (int mgrp = 0; mgrp < B; ++mgrp)
                for (int s_ = 0; s_ < p; ++s_) {
                    if (s_ == me) dst_off[(size_t)b * B + mgrp] = run;
                    run += count_of(s_, b * B + mgrp);
                }
            if (run != blk.size(b)) throw std::string("sharded construction: exchange plan does not cover the block");
        }
        for (int kk = 0; kk < p; ++kk) {
            const int b = (me + kk) % p;
            cudaStream_t s2 = kk == 0 ? e->copy_streams[0] : e->copy_streams[NCS > 1 ? 1 + (kk - 1) % (NCS - 1) : 0];
            for (int mgrp = 0; mgrp < B; ++mgrp) {
                const int bin = b * B + mgrp;
                const u64 len = count_of(me, bin);
                if (len)
                    PSAC_CUDA(cudaMemcpyAsync(at(b, oR0) + dst_off[(size_t)b * B + mgrp], Sbuf + h_bstart[(size_t)me * RADIX + bin], len * sizeof(u64),
                                              cudaMemcpyDeviceToDevice, s2));
            }
        }
        for (int i = 1; i < NCS; ++i) {
            PSAC_CUDA(cudaEventRecord(e->ev_cs[i], e->copy_streams[i]));
            PSAC_CUDA(cudaStreamWaitEvent(cs, e->ev_cs[i], 0));
        }
        PSAC_CUDA(cudaEventRecord(e->ev_xt[1], cs));
        // all pushes have landed everywhere: barrier on the copy stream, then the main stream may read the receive buffer
        ShardComm C2{reinterpret_cast<ncclComm_t>(e->nccl_comm2 ? e->nccl_comm2 : e->nccl_comm), C.rank, C.world};
        u64* d_b = e->shard_meta() + 60;
        PSAC_NCCL(g_nccl.AllReduce(d_b, d_b, 1, ncclUint64, ncclSum, C2.comm, cs));
        PSAC_CUDA(cudaEventRecord(e->ev_x[1], cs));
        PSAC_CUDA(cudaEventRecord(e->ev_xt[2], cs));
        e->xt_used = true;
    }
    // unresolved: mine and the total over the ranks
    u64* d_m = e->shard_meta() + 32;  // [0] mine, [1] sum
    auto read_unresolved = [&](u64* mine, u64* total) {
        PSAC_CUDA(cudaMemcpyAsync(d_m, e->counts(), sizeof(u64), cudaMemcpyDeviceToDevice, st));
        PSAC_NCCL(g_nccl.AllReduce(d_m, d_m + 1, 1, ncclUint64, ncclSum, C.comm, st));
        PSAC_CUDA(cudaMemcpyAsync(e->h_pinned, d_m, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        PSAC_CUDA(cudaStreamSynchronize(st));
        *mine = e->h_pinned[0];
        *total = e->h_pinned[1];
    };
    u64 m = 0, M = 0;
    read_unresolved(&m, &M);
    S.unresolved_after_first = M;
    S.rounds = 1;
    if (m > ucap) {
        // repetitive text: the list overflowed its first allocation -- list again into a large enough one
        ucap = m;
        reserve_lists(1, ucap);
        H.pos_out = e->rp[1].p;
        H.head_out = e->rh[1].as<u8>();
        H.suf_out = e->rv[1].p;
        H.cap = ucap;
        heads_kernel<u64, u64, 1, true><<<(unsigned)htiles, HD_THREADS, 0, st>>>(H);
        e->launches += 1;
        PSAC_CUDA(cudaGetLastError());
    }

    e->mark("counts");
    // ---- SA -> ISA, parts B and S: once all runs have landed, the receiver's partition pass by the top index bits and the
    //      windowed scatter into my ISA block
    {
        PSAC_CUDA(cudaStreamWaitEvent(st, e->ev_x[1], 0));
        e->mark("isa_landed");
        const u64* words = at(me, oR0);
        if (n_local >= (1ull << 22)) {
            ArraySrc<u64, NoVal> wsrc{words, nullptr, nullptr, idx_shift + hshift, (u32)(RADIX - 1), 0ull};
            launch_pass<ArraySrc<u64, NoVal>, NoVal, false>(ws, wsrc, at(me, oR1), nullptr, nullptr, n_local, st);
            e->launches += LAUNCHES_PER_PASS;
            words = at(me, oR1);
        }
        e->mark("isa_window");
        PackedScatterArgs PA{};
        PA.words = words;
        PA.isa = ISA;
        PA.n = n_local;
        PA.rank_shift = rank_shift;
        PA.idx_shift = idx_shift;
        PA.rank_mask = (1u << rank_bits) - 1u;
        PA.rel_mask = (1ull << rel_bits) - 1ull;
        for (int r = 0; r < 16; ++r) PA.off_key[r] = r < p ? P.off_key[r] : 0;
        if (n_local) {
            isa_scatter_packed_kernel<<<(unsigned)div_up(n_local, (size_t)4096), 256, 0, st>>>(PA);
            e->launches += 1;
        }
        PSAC_CUDA(cudaGetLastError());
    }
    e->mark("isa_scatter");
    e->end(PH_ISA);

    // ---- later rounds, distributed: every rank sorts its own unresolved suffixes
    if (M > 0) {
        e->begin(PH_ROUNDS);
        PeerIsa pisa{};
        for (int r = 0; r < p; ++r) pisa.blk[r] = at(r, oISA);
        pisa.div = BlkDiv::make(n, p);
        pisa.p = p;
        const int kb = (int)bits_for(n);
        auto round_args = [&](const void* pos_in, const u8* head_in, const void* suf_in, u64 mm, u64 h) {
            RoundKeyArgs Kk{};
            Kk.pos = pos_in;
            Kk.head = head_in;
            Kk.sa = nullptr;
            Kk.isa = nullptr;
            Kk.m = mm;
            Kk.n = n;
            Kk.h = h;
            Kk.kbits = kb;
            Kk.keys = e->rk[0].as<u64>();
            Kk.vals = e->vals2.p;
            Kk.lb_max = e->lookback.as<u64>();
            Kk.tile_counter = e->counters() + 17;
            Kk.suf_in = suf_in;
            Kk.rank2 = nullptr;
            Kk.pisa = pisa;
            Kk.pos_add = off;
            return Kk;
        };
        const void* pos_in = e->rp[1].p;
        const u8* head_in = e->rh[1].as<u8>();
        const void* suf_in = e->rv[1].p;
        if (m > 0) {
            e->rk[0].reserve(m * sizeof(u64), tot);
            e->rk[1].reserve(m * sizeof(u64), tot);
            e->vals2.reserve(m * sizeof(u64), tot);
            e->scratch.reserve(m * sizeof(u64), tot);
            reserve_lists(0, m);
            e->lookback.reserve(std::max<size_t>(lookback_bytes(m), e->lookback.cap), tot);
        }
        rank_barrier(e, C);  // every rank's ISA scatter is complete
        if (m > 0) {
            // the ISA entries of the unresolved suffixes get the position of their bucket head
            const u64 ntiles = div_up(m, (size_t)RES_TILE);
            RoundKeyArgs Kf = round_args(pos_in, head_in, suf_in, m, 0);
            PSAC_CUDA(cudaMemsetAsync(Kf.lb_max, 0, ntiles * sizeof(u64), st));
            PSAC_CUDA(cudaMemsetAsync(Kf.tile_counter, 0, sizeof(u32), st));
            round_keys_kernel<u64, true><<<(unsigned)ntiles, RES_THREADS, 0, st>>>(Kf);
            e->launches += 1;
            PSAC_CUDA(cudaGetLastError());
        }
        rank_barrier(e, C);
        u64 h = Cc;
        int t = 0;
        while (M > 0) {
            const u64 ntiles = div_up(m ? m : 1, (size_t)RES_TILE);
            const int mbits = m ? (int)bits_for(m - 1) : 0;
            if (kb + mbits > 64) throw arg_failure{"text too repetitive for a 64-bit round key (more than 2^30 unresolved suffixes on one rank of a text beyond 2^32)"};
            if (m > 0) {
                RoundKeyArgs Kk = round_args(pos_in, head_in, suf_in, m, h);
                PSAC_CUDA(cudaMemsetAsync(Kk.lb_max, 0, ntiles * sizeof(u64), st));
                PSAC_CUDA(cudaMemsetAsync(Kk.tile_counter, 0, sizeof(u32), st));
                round_keys_kernel<u64, false><<<(unsigned)ntiles, RES_THREADS, 0, st>>>(Kk);  // bulk get of ISA[s + h] over peer memory
                e->launches += 1;
                PSAC_CUDA(cudaGetLastError());
            }
            rank_barrier(e, C);  // all reads of this round's ISA state are done before anybody writes the next one
            if (m > 0) {
                uint64_t sl2 = 0;
                u64* vals0 = e->vals2.as<u64>();
                u64* vals1 = e->scratch.as<u64>();
                const bool a2 = radix_sort_pairs<u64, u64>(e->radix_ws(), e->rk[0].as<u64>(), e->rk[1].as<u64>(), vals0, vals1, m, 0, kb + mbits, st, e->sm_count,
                                                           nullptr, &sl2);
                e->launches += sl2;
                ResolveArgs Q{};
                Q.keys = e->rk[a2 ? 1 : 0].p;
                Q.vals = a2 ? vals1 : vals0;
                Q.pos_in = pos_in;
                Q.m = m;
                Q.n = n;
                Q.sa = Wx;  // SA[pos] = suffix (the word's key bits are spent)
                Q.isa = nullptr;
                Q.bucket_out = nullptr;
                Q.lcp = LCP;
                Q.pos_out = e->rp[t].p;
                Q.head_out = e->rh[t].as<u8>();
                Q.suf_out = e->rv[t].p;
                Q.cap = m;
                Q.counts = e->counts();
                Q.lb_max = e->lookback.as<u64>();
                Q.lb_sum = Q.lb_max + ntiles;
                Q.stream = stream;
                Q.lbits = lbits;
                Q.C = (int)Cc;
                Q.kbits = kb;
                Q.h = h;
                Q.padded_lcp = 0;
                Q.pos_base = 0;
                Q.halo = nullptr;
                Q.sa_lo = 0;
                Q.sa_hi = cnt;
                Q.isa_lo = 0;
                Q.isa_hi = 0;
                Q.pisa = pisa;
                Q.isa_add = off;
                Q.widx = widx;
                Q.lcp_main = lcp_main;
                Q.main_lo = main_lo;
                Q.main_hi = main_hi;
                launch_resolve<u64, u64>(e, false, Q);  // new bucket ids go to the owners' ISA blocks over peer memory
            } else {
                PSAC_CUDA(cudaMemsetAsync(e->counts(), 0, 2 * sizeof(u64), st));
            }
            read_unresolved(&m, &M);  // (the all-reduce also orders this round's ISA writes before the next round's reads)
            pos_in = e->rp[t].p;
            head_in = e->rh[t].as<u8>();
            suf_in = e->rv[t].p;
            t ^= 1;
            h *= 2;
            S.rounds += 1;
            if (h > 4 * n + 64 && M > 0) throw std::string("prefix doubling did not converge");
        }
        e->end(PH_ROUNDS);
    }

    e->mark("rounds");
    // ---- outputs: SA pulled from the key-range owners into exact blocks; LCP re-balanced; ISA is block-distributed already
    e->begin(PH_OUTPUT);
    {
        rank_barrier(e, C);  // every rank's words are final
        if (n_local) {
            PullArgs PL{};
            for (int r = 0; r < p; ++r) PL.src[r] = at(r, oW[x]);
            for (int r = 0; r <= p; ++r) PL.off_key[r] = P.off_key[r];
            PL.p = p;
            PL.dst_lo = blk.start(me);
            PL.m = n_local;
            PL.widx = widx;
            PL.dst = sa_out;
            if (index_bytes == 8)
                pull_sa_kernel<u64><<<grid_for(e, n_local, 256, 16), 256, 0, st>>>(PL);
            else
                pull_sa_kernel<u32><<<grid_for(e, n_local, 256, 16), 256, 0, st>>>(PL);
            e->launches += 1;
        }
        e->mark("out_sa");
        const u64 text_lo = blk.start(me), text_hi = text_lo + n_local;
        if (want_lcp) {
            std::vector<u64> scount(p), sdispl(p), rcount(p), rdispl(p);
            for (int b = 0; b < p; ++b) {
                const u64 lo = std::max(off, blk.start(b)), hi = std::min(off + cnt, blk.start(b) + blk.size(b));
                scount[b] = hi > lo ? hi - lo : 0;
                sdispl[b] = hi > lo ? lo - off : 0;
                const u64 rlo = std::max(P.off_key[b], text_lo), rhi = std::min(P.off_key[b + 1], text_hi);
                rcount[b] = rhi > rlo ? rhi - rlo : 0;
                rdispl[b] = rhi > rlo ? rlo - text_lo : 0;
            }
            if (lcp_direct) scount[me] = rcount[me] = 0;  // my own share is in place already
            const bool direct = index_bytes == 8;
            u64* target = direct ? reinterpret_cast<u64*>(lcp_out) : at(me, oR0);  // (the exchange buffer is free again)
            all_to_all_v(e, C, LCP, scount, sdispl, target, rcount, rdispl, sizeof(u64));
            if (!direct && n_local) {
                convert_kernel<u64, u32><<<grid_for(e, n_local, 256, 16), 256, 0, st>>>(target, reinterpret_cast<u32*>(lcp_out), n_local);
                e->launches += 1;
            }
        }
        e->mark("out_lcp");
        if (isa_out && n_local) {
            if (index_bytes == 8)
                PSAC_CUDA(cudaMemcpyAsync(isa_out, ISA, n_local * sizeof(u64), cudaMemcpyDeviceToDevice, st));
            else {
                convert_kernel<u64, u32><<<grid_for(e, n_local, 256, 16), 256, 0, st>>>(ISA, reinterpret_cast<u32*>(isa_out), n_local);
                e->launches += 1;
            }
        }
        PSAC_CUDA(cudaGetLastError());
    }
    e->mark("out_isa");
    e->end(PH_OUTPUT);
    return true;
}

// Inputs too small to shard: all-gather the text and let every GPU build the whole index; each keeps its block.
void gather_text(psacb200_engine* e, const ShardComm& C, const u8* d_text_local, u64 n_local, u64 n) {
    const BlkDist blk(n, C.world);
    if (blk.size(C.rank) != n_local) throw arg_failure{"the input text must be equally block decomposed across all ranks (reference suffix_array.hpp:226)"};
    e->text.reserve(n + 64, &e->device_bytes);
    std::vector<u64> cnt(C.world), dsp(C.world);
    for (int r = 0; r < C.world; ++r) {
        cnt[r] = blk.size(r);
        dsp[r] = blk.start(r);
    }
    if (n_local) PSAC_CUDA(cudaMemcpyAsync(e->text.as<u8>() + dsp[C.rank], d_text_local, n_local, cudaMemcpyDeviceToDevice, e->stream));
    all_gather_v(e, C, e->text.p, cnt, dsp, 1);
}

// ------------------------------------------------------------------------------------------------ sharded certificate
// d_check_sa over the ranks (check_kernels.cuh): every rank checks its block of SA positions; the ISA blocks are copied
// into the peer arena and read by all ranks over NVLink; the element before a block comes from the previous non-empty
// rank.  The failure counts are summed over the ranks, so every rank returns the same report.
void check_sharded_core(psacb200_engine* e, const ShardComm& C, const u8* d_text_local, u64 n_local, u64 n, int index_bytes, const void* d_sa, const void* d_isa,
                        const void* d_lcp, psacb200_check_report* rep) {
    const int p = C.world, me = C.rank;
    cudaStream_t st = e->stream;
    const BlkDist blk(n, p);
    if (blk.size(me) != n_local) throw arg_failure{"the arrays must be equally block decomposed across all ranks (reference suffix_array.hpp:226)"};
    memset(rep, 0, sizeof(*rep));
    rep->n = n;
    rep->first_bad = ~0ull;
    rep->checked_lcp = d_lcp != nullptr;
    if (n == 0) return;
    Alphabet alpha;
    prepare_text_sharded(e, C, d_text_local, n_local, n, alpha);
    const size_t blk_bytes = align_up((blk.size(0) + 16) * (size_t)index_bytes, 256);
    if (!arena_ensure(e, C, blk_bytes)) throw std::string("sharded check needs peer access between the GPUs (CUDA IPC / P2P unavailable)");
    PeerArena& A = *reinterpret_cast<PeerArena*>(e->peer_map);
    cudaEvent_t e0 = e->ev_begin[PH_OUTPUT], e1 = e->ev_end[PH_OUTPUT];
    PSAC_CUDA(cudaEventRecord(e0, st));
    rank_barrier(e, C);  // nobody still reads the arena from a previous step
    if (n_local) PSAC_CUDA(cudaMemcpyAsync(A.base, d_isa, n_local * (size_t)index_bytes, cudaMemcpyDeviceToDevice, st));
    // last SA element of every rank
    u64* d_last = e->shard_meta();  // [p]
    PSAC_CUDA(cudaMemsetAsync(d_last + me, 0, sizeof(u64), st));
    if (n_local)
        PSAC_CUDA(cudaMemcpyAsync(d_last + me, reinterpret_cast<const u8*>(d_sa) + (n_local - 1) * (size_t)index_bytes, (size_t)index_bytes, cudaMemcpyDeviceToDevice, st));
    PSAC_NCCL(g_nccl.AllGather(d_last + me, d_last, 1, ncclUint64, C.comm, st));  // (also orders the ISA copies before the peers' reads)
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned, d_last, (size_t)p * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    u64 halo = 0;
    for (int r = me - 1; r >= 0; --r)
        if (blk.size(r)) {
            halo = e->h_pinned[r];
            break;
        }
    rank_barrier(e, C);
    check_reset(e);
    if (n_local) {
        CheckArgs K{};
        K.sa = d_sa;
        K.lcp = d_lcp;
        K.pos0 = blk.start(me);
        K.m = n_local;
        K.n = n;
        K.halo_sa = halo;
        K.stream = e->packed.as<u64>();
        K.lbits = alpha.lbits;
        K.padded = alpha.zero_code_used ? 1 : 0;
        K.p = p;
        for (int r = 0; r < p; ++r) K.isa_blk[r] = A.peer[r];
        K.div = BlkDiv::make(n, p);
        K.bad = check_counters(e);
        check_launch(e, K, index_bytes);
    }
    unsigned long long* bad = check_counters(e);
    PSAC_NCCL(g_nccl.AllReduce(bad, bad, 4, ncclUint64, ncclSum, C.comm, st));
    PSAC_NCCL(g_nccl.AllReduce(bad + 4, bad + 4, 1, ncclUint64, ncclMin, C.comm, st));
    PSAC_CUDA(cudaEventRecord(e1, st));
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 48, bad, 5 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    rep->bad_range = e->h_pinned[48];
    rep->bad_inverse = e->h_pinned[49];
    rep->bad_order = e->h_pinned[50];
    rep->bad_lcp = e->h_pinned[51];
    rep->first_bad = e->h_pinned[52];
    cudaEventElapsedTime(&rep->ms, e0, e1);
}

// ================================================================================================ ANSV / suffix tree, device resident
// psacb200_ansv_device / _ansv_sharded and psacb200_suffix_tree_device / _suffix_tree_sharded (tree_kernels.cuh).  `C` may be
// null (one GPU, no communicator): then the values are searched with LocalSearch and nothing leaves the GPU.
struct TreeLayout {
    size_t o_lvl0, o_upper, o_queue, total;
    u64 qcap;
};
template <typename T>
TreeLayout tree_layout(u64 n_loc_max, int p) {
    TreeLayout L;
    L.o_lvl0 = 0;
    L.o_upper = align_up((n_loc_max + 64) * sizeof(T), 256);
    const size_t upper = align_up((n_loc_max / (ANSV_FAN - 1) + 64 * ANSV_MAX_LEVELS) * sizeof(T), 256);
    L.o_queue = L.o_upper + upper;
    L.qcap = 1ull << 20;
    L.total = L.o_queue + (size_t)p * L.qcap * 3 * sizeof(u64);
    return L;
}

// levels of the min-tree over `m` values at `base + o_lvl0` (level 0 must be in place); returns the tree with pointers
// relative to `base` (which may be a peer's mapping of the same layout)
template <typename T>
MinTree<T> mintree_describe(u8* base, const TreeLayout& L, u64 m) {
    MinTree<T> t{};
    t.level[0] = reinterpret_cast<const T*>(base + L.o_lvl0);
    t.size[0] = m;
    t.levels = 1;
    T* up = reinterpret_cast<T*>(base + L.o_upper);
    while (t.size[t.levels - 1] > (u64)ANSV_FAN) {
        if (t.levels >= ANSV_MAX_LEVELS) throw arg_failure{"ANSV input too large"};
        const u64 mm = div_up(t.size[t.levels - 1], (size_t)ANSV_FAN);
        t.level[t.levels] = up;
        t.size[t.levels] = mm;
        up += mm;
        t.levels += 1;
    }
    return t;
}
template <typename T>
void mintree_build(psacb200_engine* e, const MinTree<T>& t) {
    for (int lv = 1; lv < t.levels; ++lv) {
        mintree_level_kernel<T><<<grid_for(e, t.size[lv], 256, 8), 256, 0, e->stream>>>(t.level[lv - 1], t.size[lv - 1], const_cast<T*>(t.level[lv]), t.size[lv]);
        e->launches += 1;
    }
    PSAC_CUDA(cudaGetLastError());
}

// Distributed search structure over the values `d_vals_local` (this rank's block of a block-distributed array of n values):
// copies the block into the peer arena, builds the min-tree, all-gathers the block minima, uploads the table of all
// ranks' trees.  Returns the searcher.
template <typename T>
DistSearch<T> dist_search_setup(psacb200_engine* e, const ShardComm& C, const T* d_vals_local, u64 n_local, u64 n, TreeLayout& L) {
    const int p = C.world, me = C.rank;
    cudaStream_t st = e->stream;
    const BlkDist blk(n, p);
    L = tree_layout<T>(blk.size(0), p);
    if (!arena_ensure(e, C, L.total)) throw std::string("sharded ANSV / suffix tree needs peer access between the GPUs (CUDA IPC / P2P unavailable)");
    PeerArena& A = *reinterpret_cast<PeerArena*>(e->peer_map);
    rank_barrier(e, C);  // nobody reads my arena any more (previous call)
    if (n_local) PSAC_CUDA(cudaMemcpyAsync(A.base + L.o_lvl0, d_vals_local, n_local * sizeof(T), cudaMemcpyDeviceToDevice, st));
    const MinTree<T> mine = mintree_describe<T>(A.base, L, n_local);
    u64* d_min = e->shard_meta() + 96;  // [p] block minima (as u64)
    PSAC_CUDA(cudaMemsetAsync(d_min + me, 0, sizeof(u64), st));
    if (n_local) {
        mintree_build<T>(e, mine);
        array_min_kernel<T><<<1, 32, 0, st>>>(mine.level[mine.levels - 1], mine.size[mine.levels - 1], reinterpret_cast<T*>(d_min + me));
        e->launches += 1;
    }
    PSAC_NCCL(g_nccl.AllGather(d_min + me, d_min, 1, ncclUint64, C.comm, st));  // (also: every rank's tree is complete before anybody searches it)
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 64, d_min, (size_t)p * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    static_assert(sizeof(DistTreeTable<T>) <= 8192, "table size");
    DistTreeTable<T> tab{};
    tab.p = p;
    for (int r = 0; r < p; ++r) {
        tab.t[r] = mintree_describe<T>(A.peer[r], L, blk.size(r));
        T v;
        memcpy(&v, e->h_pinned + 64 + r, sizeof(T));
        tab.blockmin[r] = v;
        tab.start[r] = blk.start(r);
    }
    tab.start[p] = n;
    e->tb[1].reserve(8192 + 256, &e->device_bytes);
    memcpy(e->h_pinned + 6144, &tab, sizeof(tab));  // (pinned staging; at most 8192 bytes from word 6144 of 8192)
    PSAC_CUDA(cudaMemcpyAsync(e->tb[1].p, e->h_pinned + 6144, sizeof(tab), cudaMemcpyHostToDevice, st));
    DistSearch<T> sr{};
    sr.D = e->tb[1].as<DistTreeTable<T>>();
    sr.div = BlkDiv::make(n, p);
    sr.n_total = n;
    return sr;
}

template <typename T>
void ansv_sharded_core(psacb200_engine* e, const ShardComm& C, const T* d_vals_local, u64 n_local, u64 n, int left_type, int right_type, u64 nonsv,
                       u64* d_left, u64* d_right) {
    const BlkDist blk(n, C.world);
    if (blk.size(C.rank) != n_local) throw arg_failure{"the values must be equally block decomposed across all ranks"};
    TreeLayout L;
    DistSearch<T> sr = dist_search_setup<T>(e, C, d_vals_local, n_local, n, L);
    if (n_local) {
        launch_ansv_tile<T, DistSearch<T>>(sr, d_vals_local, blk.start(C.rank), n_local, left_type, right_type, nonsv, d_left, d_right, ansv_list(e, n_local),
                                           e->sm_count, e->stream);
        e->launches += 1;
    }
    rank_barrier(e, C);  // my arena may be reused only after every rank is done searching it
    PSAC_CUDA(cudaStreamSynchronize(e->stream));
}

// child table of the suffix tree from this rank's blocks of SA and LCP.  C == nullptr: one GPU.
template <typename IdxT>
void suffix_tree_core(psacb200_engine* e, const ShardComm* C, const u8* d_text_local, u64 n_local, u64 n, const IdxT* d_sa, const IdxT* d_lcp, u64* d_nodes,
                      size_t nodes_len, uint32_t* sigma_out) {
    const int p = C ? C->world : 1, me = C ? C->rank : 0;
    cudaStream_t st = e->stream;
    size_t* tot = &e->device_bytes;
    const BlkDist blk(n, p);
    if (blk.size(me) != n_local) throw arg_failure{"the arrays must be equally block decomposed across all ranks (reference suffix_array.hpp:226)"};
    e->tr_n = 0;
    e->mark("begin");
    Alphabet alpha;
    if (C)
        prepare_text_sharded(e, *C, d_text_local, n_local, n, alpha);
    else
        prepare_text(e, d_text_local, n, nullptr, alpha);
    if (sigma_out) *sigma_out = alpha.sigma;
    const size_t width = (size_t)alpha.sigma + 1;
    if (nodes_len < width * n_local) throw arg_failure{"nodes buffer too small: (sigma + 1) * n_local entries are needed"};
    e->mark("text");
    TreeFusedArgs<IdxT> A{};
    A.sa = d_sa;
    A.lcp = d_lcp;
    A.g0 = blk.start(me);
    A.m = n_local;
    A.n = n;
    A.stream = e->packed.as<u64>();
    A.lbits = alpha.lbits;
    A.sigma = alpha.sigma;
    A.code_add = alpha.zero_code_used ? 0u : 1u;
    A.nodes = d_nodes;
    A.me = me;
    A.div = BlkDiv::make(n, p);
    unsigned long long* d_cur = reinterpret_cast<unsigned long long*>(e->shard_meta() + 128);  // [16] cursors, [16] overflow
    PSAC_CUDA(cudaMemsetAsync(d_cur, 0, 17 * sizeof(u64), st));
    A.q.cursor = d_cur;
    A.q.overflow = d_cur + 16;
    // (small alphabets: the tile kernel assembles its rows in shared memory and writes every row of the table itself)
    if (!tree_tile_writes_all_rows<IdxT>(n, alpha.sigma)) PSAC_CUDA(cudaMemsetAsync(d_nodes, 0, width * n_local * sizeof(u64), st));
    if (C == nullptr) {
        // one GPU: min-tree in a private buffer, nothing is queued
        TreeLayout L = tree_layout<IdxT>(n, 1);
        e->tb[0].reserve(L.o_queue + 64, tot);
        PSAC_CUDA(cudaMemcpyAsync(e->tb[0].as<u8>() + L.o_lvl0, d_lcp, n * sizeof(IdxT), cudaMemcpyDeviceToDevice, st));
        LocalSearch<IdxT> sr{mintree_describe<IdxT>(e->tb[0].as<u8>(), L, n)};
        mintree_build<IdxT>(e, sr.t);
        e->mark("mintree");
        A.q.cap = 0;
        launch_tree_tile<IdxT, LocalSearch<IdxT>>(A, sr, ansv_list(e, n), e->sm_count, st);
        e->launches += 1;
        e->mark("tree");
        PSAC_CUDA(cudaStreamSynchronize(st));
        return;
    }
    TreeLayout L;
    DistSearch<IdxT> sr = dist_search_setup<IdxT>(e, *C, d_lcp, n_local, n, L);
    PeerArena& AR = *reinterpret_cast<PeerArena*>(e->peer_map);
    e->mark("mintree");
    A.q.cap = L.qcap;
    for (int r = 0; r < 16; ++r) A.q.queue[r] = r < p ? reinterpret_cast<u64*>(AR.peer[r] + L.o_queue) + (size_t)me * L.qcap * 3 : nullptr;
    if (n_local) {
        launch_tree_tile<IdxT, DistSearch<IdxT>>(A, sr, ansv_list(e, n_local), e->sm_count, st);
        e->launches += 1;
    }
    e->mark("tree");
    // edges queued for rows of other ranks: exchange the counts, apply mine
    u64* d_all = e->shard_meta() + 160;  // [p][17]
    PSAC_CUDA(cudaMemcpyAsync(d_all + (size_t)me * 17, d_cur, 17 * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    PSAC_NCCL(g_nccl.AllGather(d_all + (size_t)me * 17, d_all, 17, ncclUint64, C->comm, st));  // (orders every rank's queue writes before the apply)
    PSAC_CUDA(cudaMemcpyAsync(e->h_pinned + 128, d_all, (size_t)p * 17 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    PSAC_CUDA(cudaStreamSynchronize(st));
    u64 overflow = 0, applied = 0;
    for (int s_ = 0; s_ < p; ++s_) {
        overflow += e->h_pinned[128 + s_ * 17 + 16];
        const u64 c = e->h_pinned[128 + s_ * 17 + me];
        if (s_ == me || c == 0) continue;
        tree_apply_edges_kernel<<<grid_for(e, c, 256, 8), 256, 0, st>>>(reinterpret_cast<const u64*>(AR.base + L.o_queue) + (size_t)s_ * L.qcap * 3, c, alpha.sigma, d_nodes);
        e->launches += 1;
        applied += c;
    }
    PSAC_CUDA(cudaGetLastError());
    e->stats.unresolved_after_first = applied;  // (reported: edges that crossed ranks)
    rank_barrier(e, *C);
    e->mark("remote_edges");
    PSAC_CUDA(cudaStreamSynchronize(st));
    if (overflow) throw std::string("suffix tree: the cross-rank edge queue overflowed (") + std::to_string(overflow) + " edges)";
}

}  // namespace
