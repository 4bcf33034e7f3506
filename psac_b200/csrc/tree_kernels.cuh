// psac-b200: all-nearest-smaller-values (ANSV) and suffix-tree construction from SA + LCP.
//
// Reference restated:
//   ansv<T, left_type, right_type, indexing>   include/ansv.hpp:2042-2051 (sequential core :47-65, modes :24-45)
//   for_each_parent / construct_suffix_tree     include/suffix_tree.hpp:43-223, 440-499
// The reference runs a sequential stack pass per rank and merges the unmatched prefix minima between ranks.  On the GPU
// every element searches a MIN-TREE instead (level k holds the minima of blocks of 32^k values): walk away from i
// inside the current block of 32, climb one level when the block is exhausted, descend into the first block whose minimum
// qualifies.  LCP arrays of real text have their nearest smaller value a few positions away, so almost every search
// ends in the first block; the worst case is 31 probes per level (7 levels at n = 2^31).  No stack, no ordering between
// threads, and the three match modes of the reference become compositions of the same directional search:
//   nearest_sm  (0): nearest j with v[j] <  v[i]
//   nearest_eq  (1): nearest j with v[j] <= v[i]
//   furthest_eq (2): with s = nearest strictly smaller: the element right after s (towards i) that is <= v[i] if it lies
//                    before i (all elements between s and i are >= v[i], so it equals v[i] and is the furthest such);
//                    otherwise the far end of s's own run of equal values (same construction one step further out).
#pragma once
#include "common.cuh"

namespace psacb200 {

// Fan-out of the block-minimum tree.  An exact search scans the rest of its block at every level it climbs and one block at
// every level it descends, one dependent load per entry: ~ FAN / 2 * 2 * log_FAN(distance) loads.  8 instead of 32 cuts that
// chain for the far matches the list kernels are left with (the tree takes n / 7 instead of n / 31 extra entries).  Measured at
// 2^29 (tree / standalone ANSV): fan-out 32: 35.0 / 36.0 ms, 8: 32.9 / 31.9 ms, 4: 33.4 / 32.6 ms.
constexpr int ANSV_FAN_LOG = 3;
constexpr int ANSV_FAN = 1 << ANSV_FAN_LOG;
constexpr int ANSV_MAX_LEVELS = 14;
constexpr u64 ANSV_NONE = ~0ull;

template <typename T>
struct MinTree {
    const T* level[ANSV_MAX_LEVELS];  // level[0] = the values
    u64 size[ANSV_MAX_LEVELS];
    int levels;
};

template <typename T>
__global__ void __launch_bounds__(256) mintree_level_kernel(const T* __restrict__ in, u64 n_in, T* __restrict__ out, u64 n_out) {
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_out; b += (u64)gridDim.x * blockDim.x) {
        const u64 lo = b * ANSV_FAN, hi = (lo + ANSV_FAN < n_in) ? lo + ANSV_FAN : n_in;
        T m = in[lo];
        for (u64 j = lo + 1; j < hi; ++j) m = in[j] < m ? in[j] : m;
        out[b] = m;
    }
}

template <typename T, bool STRICT>
__device__ __forceinline__ bool ansv_hit(T v, T x) {
    return STRICT ? (v < x) : (v <= x);
}

// nearest hit inside a[lo, hi) seen from the near side (DIR < 0: the largest index, DIR > 0: the smallest), ANSV_NONE if there
// is none.  (Fetching eight entries per round trip with independent loads was measured: the registers it takes cost the list
// kernels more occupancy than the shorter dependency chains win, 46 -> 52 ms at 2^29, profiles/r2_tree_notes.md.)
template <typename T, int DIR, bool STRICT>
__device__ __forceinline__ u64 ansv_scan(const T* __restrict__ a, u64 lo, u64 hi, T x) {
    if (DIR < 0) {
        for (u64 j = hi; j > lo;) {
            --j;
            if (ansv_hit<T, STRICT>(a[j], x)) return j;
        }
    } else {
        for (u64 j = lo; j < hi; ++j)
            if (ansv_hit<T, STRICT>(a[j], x)) return j;
    }
    return ANSV_NONE;
}

// Nearest position j on the LEFT (DIR = -1: j < i) or RIGHT (DIR = +1: j > i) of i with v[j] < x (STRICT) or <= x.
// ANSV_NONE if there is none.
template <typename T, int DIR, bool STRICT>
__device__ __noinline__ u64 ansv_search(const MinTree<T>& t, u64 i, T x) {
    int lv = 0;
    u64 p = i;  // current position at level lv; its block of 32 is [p & ~31, p | 31]
    // ---- climb: scan the rest of the current block away from p, then move to the parent
    while (true) {
        const u64 n = t.size[lv];
        const u64 blo = p & ~(u64)(ANSV_FAN - 1);
        u64 bhi = blo + ANSV_FAN;
        bhi = bhi < n ? bhi : n;
        const u64 found = DIR < 0 ? ansv_scan<T, DIR, STRICT>(t.level[lv], blo, p, x) : ansv_scan<T, DIR, STRICT>(t.level[lv], p + 1, bhi, x);
        if (found != ANSV_NONE) {
            p = found;
            break;
        }
        if (lv + 1 >= t.levels) return ANSV_NONE;
        p >>= ANSV_FAN_LOG;
        ++lv;
    }
    // ---- descend: inside block p of level lv take the child nearest to i that qualifies (the block's minimum does)
    while (lv > 0) {
        --lv;
        const u64 n = t.size[lv];
        const u64 lo = p * ANSV_FAN;
        u64 hi = lo + ANSV_FAN;
        hi = hi < n ? hi : n;
        p = ansv_scan<T, DIR, STRICT>(t.level[lv], lo, hi, x);
    }
    return p;
}

// inside block-minimum tree t, the position nearest to the entry side (DIR < 0: we come from the right, so the LARGEST
// qualifying index; DIR > 0: the smallest) -- the caller knows that the minimum of the whole array qualifies
template <typename T, int DIR, bool STRICT>
__device__ __noinline__ u64 ansv_descend_top(const MinTree<T>& t, T x) {
    int lv = t.levels - 1;
    u64 p = ansv_scan<T, DIR, STRICT>(t.level[lv], 0, t.size[lv], x);
    while (lv > 0) {
        --lv;
        const u64 n = t.size[lv];
        const u64 lo = p * ANSV_FAN;
        u64 hi = lo + ANSV_FAN;
        hi = hi < n ? hi : n;
        p = ansv_scan<T, DIR, STRICT>(t.level[lv], lo, hi, x);
    }
    return p;
}

// ---- searchers: where the values live.  LocalSearch: one array on this GPU.  DistSearch: the array is block-distributed over
// the ranks of one box (mxx::blk_dist); every rank's min-tree lives in its peer-visible arena, the minima of the blocks are
// replicated.  A search first runs in the block of its start position; only if that block is exhausted (the element is a
// prefix / suffix minimum of its block -- the reference's "unmatched" elements, ansv.hpp:1362-1442) it walks over the block
// minima and descends into the first block that qualifies, reading that rank's tree through peer memory.
template <typename T>
struct LocalSearch {
    MinTree<T> t;
    __device__ __forceinline__ u64 n() const { return t.size[0]; }
    __device__ __forceinline__ T value(u64 g) const { return t.level[0][g]; }
    template <int DIR, bool STRICT>
    __device__ __forceinline__ u64 search(u64 g, T x) const {
        return ansv_search<T, DIR, STRICT>(t, g, x);
    }
};

template <typename T>
struct DistTreeTable {  // lives in device memory
    MinTree<T> t[16];
    T blockmin[16];
    u64 start[17];
    int p;
};

template <typename T>
struct DistSearch {
    const DistTreeTable<T>* D;
    BlkDiv div;
    u64 n_total;
    __device__ __forceinline__ u64 n() const { return n_total; }
    __device__ __forceinline__ T value(u64 g) const {
        u64 l;
        const u32 r = div.owner(g, &l);
        return D->t[r].level[0][l];
    }
    template <int DIR, bool STRICT>
    __device__ u64 search(u64 g, T x) const {
        u64 l;
        const int r = (int)div.owner(g, &l);
        const u64 res = ansv_search<T, DIR, STRICT>(D->t[r], l, x);
        if (res != ANSV_NONE) return D->start[r] + res;
        for (int rr = r + DIR; rr >= 0 && rr < D->p; rr += DIR) {
            if (D->t[rr].size[0] == 0) continue;
            if (ansv_hit<T, STRICT>(D->blockmin[rr], x)) return D->start[rr] + ansv_descend_top<T, DIR, STRICT>(D->t[rr], x);
        }
        return ANSV_NONE;
    }
};

// one side of the ANSV of element i (global position) under the reference's match mode
template <typename T, int DIR, class S>
__device__ u64 ansv_one(const S& sr, u64 i, int mode) {
    const T x = sr.value(i);
    if (mode == 1) return sr.template search<DIR, false>(i, x);
    const u64 s = sr.template search<DIR, true>(i, x);
    if (mode == 0) return s;
    // furthest_eq.  Everything between s (or the array end if there is no s) and i is >= x, so walking from there back
    // towards i the first element <= x equals x and is the furthest equal one; the walk stops at i itself at the latest.
    const u64 n = sr.n();
    const u64 end = DIR < 0 ? 0 : n - 1;
    u64 e;
    if (s != ANSV_NONE)
        e = sr.template search<-DIR, false>(s, x);
    else if (end == i)
        e = i;
    else
        e = (sr.value(end) <= x) ? end : sr.template search<-DIR, false>(end, x);
    if (e != i) return e;
    if (s == ANSV_NONE) return ANSV_NONE;
    // no equal element before s: the far end of s's own run of equal values (same construction one step further out)
    const T m = sr.value(s);
    const u64 s2 = sr.template search<DIR, true>(s, m);
    if (s2 != ANSV_NONE) return sr.template search<-DIR, false>(s2, m);
    if (end == s) return s;
    return (sr.value(end) <= m) ? end : sr.template search<-DIR, false>(end, m);
}

// left / right: m entries for the global positions [g0, g0 + m)
template <typename T, class S>
__global__ void __launch_bounds__(256) ansv_kernel(S sr, u64 g0, u64 m, int left_mode, int right_mode, u64 nonsv, u64* __restrict__ left,
                                                   u64* __restrict__ right) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 l = ansv_one<T, -1>(sr, g0 + i, left_mode);
        const u64 r = ansv_one<T, +1>(sr, g0 + i, right_mode);
        left[i] = l == ANSV_NONE ? nonsv : l;
        right[i] = r == ANSV_NONE ? nonsv : r;
    }
}

// ------------------------------------------------------------------ suffix tree (child table)
// nodes[(sigma+1) * parent + code(first edge character)] = child id; internal node ids are LCP indices, leaf ids are
// n + SA position, 0 = empty; a child whose edge starts past the end of the text goes to column 0
// (reference suffix_tree.hpp:466-496).  One thread per SA position handles its leaf edge and its internal-node edge.
struct TreeArgs {
    const void* sa;
    const void* lcp;
    const u8* text;
    u64 n;
    const u64* left;   // ANSV of the LCP array: furthest_eq to the left, nearest_sm to the right (suffix_tree.hpp:62)
    const u64* right;
    u32 sigma;
    u8 lut[256];       // reference alphabet codes (alphabet.hpp:157-164)
    u64* nodes;
};

template <typename IdxT>
__device__ __forceinline__ void tree_emit(const TreeArgs& A, u64 parent, u64 gidx, u64 sa_val, u64 lcp_val) {
    const u64 ci = sa_val + lcp_val;
    const u64 col = ci < A.n ? (u64)A.lut[A.text[ci]] : 0;
    A.nodes[parent * (u64)(A.sigma + 1) + col] = gidx;
}

template <typename IdxT>
__global__ void __launch_bounds__(256) suffix_tree_kernel(TreeArgs A) {
    const IdxT* SA = reinterpret_cast<const IdxT*>(A.sa);
    const IdxT* LCP = reinterpret_cast<const IdxT*>(A.lcp);
    const u64 n = A.n;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 lcp_i = LCP[i];
        // ---- leaf n + i (suffix_tree.hpp:88-141)
        u64 parent, lcp_val;
        if (i == 0) {
            lcp_val = n > 1 ? (u64)LCP[1] : 0;
            parent = lcp_val > 0 ? 1 : 0;
        } else if (i == n - 1 || lcp_i >= (u64)LCP[i + 1]) {
            const u64 nsv = A.left[i];
            lcp_val = LCP[nsv];
            if (lcp_val == lcp_i) {
                parent = nsv;
            } else {
                parent = i;
                lcp_val = lcp_i;
            }
        } else {
            parent = i + 1;
            lcp_val = LCP[i + 1];
        }
        tree_emit<IdxT>(A, parent, n + i, SA[i], lcp_val);
        // ---- internal node i (suffix_tree.hpp:146-222): the root (i = 0) and its duplicates (LCP = 0) have no parent
        if (i == 0 || lcp_i == 0) continue;
        const u64 lnsv = A.left[i];
        const u64 left_val = LCP[lnsv];
        if (A.right[i] == ANSV_NONE) {
            if (left_val == lcp_i) continue;  // duplicate of the node further left
            tree_emit<IdxT>(A, lnsv, i, SA[i], left_val);
        } else {
            const u64 rnsv = A.right[i];
            const u64 right_val = LCP[rnsv];
            if (left_val >= right_val) {
                if (left_val == lcp_i) continue;
                tree_emit<IdxT>(A, lnsv, i, SA[i], left_val);
            } else {
                tree_emit<IdxT>(A, rnsv, i, SA[i], right_val);
            }
        }
    }
}

// ------------------------------------------------------------------ suffix tree, device resident and sharded
// The same child table, built from DEVICE blocks of SA and LCP without materialising the ANSV arrays: every position
// searches its left (furthest_eq) and right (nearest_sm) match on the fly (reference for_each_parent, suffix_tree.hpp:43-223,
// which calls ansv<index_t, furthest_eq, nearest_sm, local_indexing> at :62).  Rank r owns the table rows of the LCP indices
// of its block; an edge whose parent row lives on another rank is appended to that rank's edge queue through peer memory and
// applied by tree_apply_edges_kernel after a barrier (reference :440-499 sends these edges with an all-to-all).
struct TreeQueues {
    u64* queue[16];        // queue[dst]: my sub-queue inside rank dst's arena, cap entries of {row, column, child}
    unsigned long long* cursor;  // [16] entries written so far per destination (local memory)
    u64 cap;
    unsigned long long* overflow;
};

template <typename IdxT>
struct TreeFusedArgs {
    const IdxT* sa;        // local block
    const IdxT* lcp;       // local block
    u64 g0, m, n;          // first global position, local size, text length
    const u64* stream;     // packed text (replicated)
    int lbits;
    u32 sigma, code_add;   // column of a character = dense code + code_add (reference code, alphabet.hpp:157-164)
    u64* nodes;            // my rows: (sigma + 1) * m
    int me;
    BlkDiv div;
    TreeQueues q;
};

template <typename IdxT>
__device__ __forceinline__ void tree_emit_dist(const TreeFusedArgs<IdxT>& A, u64 parent, u64 child, u64 sa_val, u64 lcp_val) {
    const u64 ci = sa_val + lcp_val;
    const u64 col = ci < A.n ? stream_extract(A.stream, ci, A.lbits, A.lbits) + A.code_add : 0;
    u64 row;
    const int r = (int)A.div.owner(parent, &row);
    if (r == A.me) {
        A.nodes[row * (u64)(A.sigma + 1) + col] = child;
    } else {
        const u64 slot = atomicAdd(&A.q.cursor[r], 1ull);
        if (slot < A.q.cap) {
            u64* e = A.q.queue[r] + 3 * slot;
            e[0] = row;
            e[1] = col;
            e[2] = child;
        } else {
            atomicAdd(A.q.overflow, 1ull);
        }
    }
}

// one position: leaf edge and internal-node edge, every search through the searcher (exact, any distance)
template <typename IdxT, class S>
__device__ void tree_element_slow(const TreeFusedArgs<IdxT>& A, const S& sr, u64 i) {
    const u64 n = A.n;
    const u64 gi = A.g0 + i;
    const u64 lcp_i = A.lcp[i];
    const u64 sa_i = A.sa[i];
    // ---- leaf n + gi (suffix_tree.hpp:88-141)
    u64 parent, lcp_val;
    u64 lnsv = ANSV_NONE, left_val = 0;
    if (gi > 0) {
        lnsv = ansv_one<IdxT, -1>(sr, gi, 2);
        left_val = sr.value(lnsv);
    }
    if (gi == 0) {
        lcp_val = n > 1 ? (u64)sr.value(1) : 0;
        parent = lcp_val > 0 ? 1 : 0;
    } else {
        const u64 next = gi + 1 < n ? (i + 1 < A.m ? (u64)A.lcp[i + 1] : (u64)sr.value(gi + 1)) : 0;
        if (gi == n - 1 || lcp_i >= next) {
            if (left_val == lcp_i) {
                parent = lnsv;
                lcp_val = left_val;
            } else {
                parent = gi;
                lcp_val = lcp_i;
            }
        } else {
            parent = gi + 1;
            lcp_val = next;
        }
    }
    tree_emit_dist<IdxT>(A, parent, n + gi, sa_i, lcp_val);
    // ---- internal node gi (suffix_tree.hpp:146-222): the root (0) and its duplicates (LCP = 0) have no parent
    if (gi == 0 || lcp_i == 0) return;
    const u64 rnsv = ansv_one<IdxT, +1>(sr, gi, 0);
    if (rnsv == ANSV_NONE) {
        if (left_val == lcp_i) return;  // duplicate of the node further left
        tree_emit_dist<IdxT>(A, lnsv, gi, sa_i, left_val);
    } else {
        const u64 right_val = sr.value(rnsv);
        if (left_val >= right_val) {
            if (left_val == lcp_i) return;
            tree_emit_dist<IdxT>(A, lnsv, gi, sa_i, left_val);
        } else {
            tree_emit_dist<IdxT>(A, rnsv, gi, sa_i, right_val);
        }
    }
}

// ------------------------------------------------------------------ tile kernels: sparse table in shared memory
// A min-tree search costs one dependent probe per step and a warp waits for its slowest lane: on the LCP array of random DNA
// a third of the positions need 20+ probes and a few per warp need hundreds (measured: 94 ms for the ANSV of 2^29 values,
// 105 ms for the tree, profiles/r2_tree_notes.md).  The tile kernels take TILE consecutive values into shared memory, build
// the sparse table M[k][j] = min(v[j .. j + 2^k)) there, and answer every nearest-smaller search INSIDE the tile by
// binary lifting: log2(TILE) branch-free steps, the same for every lane.  Only positions whose match lies outside the tile
// (a few per cent) are put on a list and finished by the exact min-tree search (ansv_one / tree_element_slow) with all
// lanes busy.  The three match modes are the same compositions of directional searches as in ansv_one.
template <typename T>
struct AnsvTile {
    static constexpr int TILE = sizeof(T) == 4 ? 2048 : 1024;
    static constexpr int K = sizeof(T) == 4 ? 11 : 10;  // log2(TILE)
    static constexpr size_t SMEM = (size_t)K * TILE * sizeof(T);
    static constexpr int NOT_IN_TILE = -1;
    const T* M;  // [K][TILE] in shared memory
    __device__ __forceinline__ T at(int k, int j) const { return M[k * TILE + j]; }
    // nearest j < i (DIR < 0) / j > i (DIR > 0) inside the tile with v[j] < x (STRICT) or v[j] <= x; NOT_IN_TILE if none
    template <int DIR, bool STRICT>
    __device__ __forceinline__ int search(int i, T x) const {
        if (DIR < 0) {
            int pos = i;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) {
                const int len = 1 << k;
                if (pos >= len) {
                    const T m = at(k, pos - len);
                    if (STRICT ? (m >= x) : (m > x)) pos -= len;  // nothing in [pos - len, pos) qualifies: skip it
                }
            }
            return pos > 0 ? pos - 1 : NOT_IN_TILE;
        } else {
            int pos = i + 1;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) {
                const int len = 1 << k;
                if (pos + len <= TILE) {
                    const T m = at(k, pos);
                    if (STRICT ? (m >= x) : (m > x)) pos += len;
                }
            }
            return pos < TILE ? pos : NOT_IN_TILE;
        }
    }
    // one side of the ANSV of tile element i under the reference's match mode; NOT_IN_TILE = the exact search must decide
    template <int DIR>
    __device__ __forceinline__ int one(int i, int mode) const {
        const T x = at(0, i);
        if (mode == 1) return search<DIR, false>(i, x);
        const int s = search<DIR, true>(i, x);
        if (mode == 0 || s == NOT_IN_TILE) return s;
        const int e = search<-DIR, false>(s, x);  // walking back from s towards i: the first element <= x (i itself at the latest)
        if (e != i) return e;
        const T m = at(0, s);  // no equal element before s: the far end of s's own run of equal values
        const int s2 = search<DIR, true>(s, m);
        if (s2 == NOT_IN_TILE) return NOT_IN_TILE;
        return search<-DIR, false>(s2, m);
    }
};

// loads the tile [t0, t0 + TILE) of vals (m values; past the end: the largest value, so that searches run off the tile) and
// builds the sparse table
template <typename T, typename V>
__device__ __forceinline__ void ansv_tile_build(V* M, const T* __restrict__ vals, u64 t0, u64 m) {
    constexpr int TILE = AnsvTile<V>::TILE, K = AnsvTile<V>::K;
    for (int j = threadIdx.x; j < TILE; j += blockDim.x) M[j] = (t0 + j < m) ? (V)vals[t0 + j] : (V)~(V)0;
    __syncthreads();
    for (int k = 1; k < K; ++k) {
        const int half = 1 << (k - 1);
        for (int j = threadIdx.x; j + 2 * half <= TILE; j += blockDim.x) {
            const V a = M[(k - 1) * TILE + j], b = M[(k - 1) * TILE + j + half];
            M[k * TILE + j] = a < b ? a : b;
        }
        __syncthreads();
    }
}

// Positions whose match lies outside their tile are appended to a GLOBAL list (entry = local index << 2 | left flag | right flag)
// and finished by a second kernel, ansv_list_kernel / tree_list_kernel, where every thread has one of them: the exact search
// is a chain of dependent loads, and only a full grid of such chains hides their latency (finishing them inside the tile
// kernel, a few dozen per tile, left most of the CTA waiting: 170 ms for 2^29 positions, profiles/r2_tree_notes.md).
struct AnsvList {
    u64* entries;
    unsigned long long* count;
    u64 cap;
};

template <typename T, class S>
__global__ void __launch_bounds__(256, 2) ansv_tile_kernel(S sr, const T* __restrict__ vals, u64 g0, u64 m, int left_mode, int right_mode, u64 nonsv,
                                                          u64* __restrict__ left, u64* __restrict__ right, AnsvList L) {
    extern __shared__ __align__(16) unsigned char ansv_smem[];
    T* M = reinterpret_cast<T*>(ansv_smem);
    using Tile = AnsvTile<T>;
    const u64 t0 = (u64)blockIdx.x * Tile::TILE;
    ansv_tile_build<T, T>(M, vals, t0, m);
    const Tile tile{M};
    for (int j = threadIdx.x; j < Tile::TILE && t0 + j < m; j += blockDim.x) {
        const int l = tile.template one<-1>(j, left_mode), r = tile.template one<+1>(j, right_mode);
        // a match that is the padding past the end of the array is no match: let the exact search decide
        const bool lf = l == Tile::NOT_IN_TILE, rf = r == Tile::NOT_IN_TILE || t0 + (u64)r >= m;
        if (!lf) left[t0 + j] = g0 + t0 + (u64)l;
        if (!rf) right[t0 + j] = g0 + t0 + (u64)r;
        if (lf || rf) {
            const u64 slot = atomicAdd(L.count, 1ull);  // (the list has room for every position)
            if (slot < L.cap) L.entries[slot] = ((t0 + (u64)j) << 2) | (lf ? 1u : 0u) | (rf ? 2u : 0u);
        }
    }
}

template <typename T, class S>
__global__ void __launch_bounds__(256) ansv_list_kernel(S sr, u64 g0, int left_mode, int right_mode, u64 nonsv, u64* __restrict__ left, u64* __restrict__ right,
                                                        AnsvList L) {
    const u64 cnt = *L.count < L.cap ? *L.count : L.cap;
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (u64)gridDim.x * blockDim.x) {
        const u64 entry = L.entries[e], i = entry >> 2;
        if (entry & 1u) {
            const u64 v = ansv_one<T, -1>(sr, g0 + i, left_mode);
            left[i] = v == ANSV_NONE ? nonsv : v;
        }
        if (entry & 2u) {
            const u64 v = ansv_one<T, +1>(sr, g0 + i, right_mode);
            right[i] = v == ANSV_NONE ? nonsv : v;
        }
    }
}

// the fused child-table fill, tile version: positions whose left (furthest_eq) and right (nearest_sm) matches and whose
// successor lie inside the tile are finished from shared memory; the others go on the list (tree_list_kernel)
// (V = type of the values in shared memory: 32 bits whenever the text is shorter than 2^32 -- LCP values are below n --, which
//  doubles the tile)
// STAGE: the child-table rows of the tile are assembled in shared memory and leave as one coalesced stream.  A position that
// is finished here has its matches -- hence its parent -- inside the tile, so every edge the tile kernel emits lands in the
// tile's own rows; the table needs no clearing pass, and the scattered 8-byte stores into 40-byte rows (a read-modify-write of
// a DRAM sector each) disappear.  Needs (sigma + 1) * TILE * 8 bytes next to the sparse table: small alphabets (DNA).
template <typename IdxT, typename V, class S, bool STAGE>
__global__ void __launch_bounds__(STAGE ? 1024 : 256, STAGE ? 1 : 2) suffix_tree_tile_kernel(TreeFusedArgs<IdxT> A, S sr, AnsvList L) {
    extern __shared__ __align__(16) unsigned char ansv_smem[];
    V* M = reinterpret_cast<V*>(ansv_smem);
    using Tile = AnsvTile<V>;
    const u64 t0 = (u64)blockIdx.x * Tile::TILE;
    u64* rows = reinterpret_cast<u64*>(ansv_smem + Tile::SMEM);  // [TILE][sigma + 1] (STAGE)
    const u32 width = A.sigma + 1;
    if (STAGE) {
        ulonglong2* rz = reinterpret_cast<ulonglong2*>(rows);
        for (u32 e = threadIdx.x; e < (u32)Tile::TILE * width / 2; e += blockDim.x) rz[e] = make_ulonglong2(0ull, 0ull);
    }
    ansv_tile_build<IdxT, V>(M, A.lcp, t0, A.m);
    const Tile tile{M};
    const u64 n = A.n;
    // left match (furthest_eq) in two steps, so that no position repeats the work of another: (1) s = nearest strictly smaller,
    // e = first entry <= x walking back from s: an equal entry before j, or j itself; (2) where e is j itself the answer is the
    // far end of s's run of equal values -- which is what step (1) of position s has found.
    __shared__ int s_feq[Tile::TILE];
    constexpr int SELF = 1 << 30;  // "no equal entry before me; my strictly smaller one is at (value & ~SELF)"
    for (int j = threadIdx.x; j < Tile::TILE; j += blockDim.x) {
        int f = Tile::NOT_IN_TILE;
        if (t0 + j < A.m) {
            const V x = tile.at(0, j);
            const int s = tile.template search<-1, true>(j, x);
            if (s != Tile::NOT_IN_TILE) {
                const int e = tile.template search<+1, false>(s, x);
                f = e != j ? e : (SELF | s);
            }
        }
        s_feq[j] = f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Tile::TILE && t0 + j < A.m; j += blockDim.x) {
        const u64 i = t0 + j, gi = A.g0 + i;
        int l = s_feq[j];
        if (l != Tile::NOT_IN_TILE && (l & SELF)) {
            const int s = l & ~SELF, fs = s_feq[s];
            l = fs == Tile::NOT_IN_TILE ? Tile::NOT_IN_TILE : ((fs & SELF) ? s : fs);
        }
        const int r = tile.template one<+1>(j, 0);
        const bool in_tile = l != Tile::NOT_IN_TILE && r != Tile::NOT_IN_TILE && t0 + (u64)r < A.m && j + 1 < Tile::TILE && i + 1 < A.m && gi > 0;
        if (!in_tile) {
            const u64 slot = atomicAdd(L.count, 1ull);  // (the list has room for every position)
            if (slot < L.cap) L.entries[slot] = i;
            continue;
        }
        const u64 lcp_i = tile.at(0, j), sa_i = A.sa[i];
        const u64 lnsv = A.g0 + t0 + (u64)l, left_val = tile.at(0, l);
        const u64 rnsv = A.g0 + t0 + (u64)r, right_val = tile.at(0, r);
        const u64 next = tile.at(0, j + 1);
        // ---- leaf n + gi (suffix_tree.hpp:88-141; gi > 0 and gi + 1 < n here)
        u64 parent, lcp_val;
        if (lcp_i >= next) {
            if (left_val == lcp_i) {
                parent = lnsv;
                lcp_val = left_val;
            } else {
                parent = gi;
                lcp_val = lcp_i;
            }
        } else {
            parent = gi + 1;
            lcp_val = next;
        }
        // (STAGE: parent - (g0 + t0) is the parent's row inside the tile: l, j, j + 1 or r)
        auto emit = [&](u64 par, u64 child, u64 lv) {
            if (STAGE) {
                const u64 ci = sa_i + lv;
                const u64 col = ci < n ? stream_extract(A.stream, ci, A.lbits, A.lbits) + A.code_add : 0;
                rows[(par - (A.g0 + t0)) * width + col] = child;
            } else {
                tree_emit_dist<IdxT>(A, par, child, sa_i, lv);
            }
        };
        emit(parent, n + gi, lcp_val);
        // ---- internal node gi (suffix_tree.hpp:146-222)
        if (lcp_i == 0) continue;
        if (left_val >= right_val) {
            if (left_val == lcp_i) continue;
            emit(lnsv, gi, left_val);
        } else {
            emit(rnsv, gi, right_val);
        }
    }
    if (STAGE) {
        __syncthreads();
        const u64 valid_rows = A.m - t0 < (u64)Tile::TILE ? A.m - t0 : (u64)Tile::TILE;
        const u64 cells = valid_rows * width;
        u64* out = A.nodes + t0 * (u64)width;  // (t0 is a multiple of the tile: 16-byte aligned whenever the table is)
        if ((reinterpret_cast<size_t>(out) & 15) == 0) {
            const ulonglong2* src2 = reinterpret_cast<const ulonglong2*>(rows);
            ulonglong2* out2 = reinterpret_cast<ulonglong2*>(out);
            for (u64 e = threadIdx.x; e < cells / 2; e += blockDim.x) __stcs(out2 + e, src2[e]);
            if ((cells & 1) && threadIdx.x == 0) out[cells - 1] = rows[cells - 1];
        } else {
            for (u64 e = threadIdx.x; e < cells; e += blockDim.x) out[e] = rows[e];
        }
    }
}

template <typename IdxT, class S>
__global__ void __launch_bounds__(256) tree_list_kernel(TreeFusedArgs<IdxT> A, S sr, AnsvList L) {
    const u64 cnt = *L.count < L.cap ? *L.count : L.cap;
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (u64)gridDim.x * blockDim.x) tree_element_slow<IdxT, S>(A, sr, L.entries[e]);
}

// edges that other ranks queued for my rows: queue of source s holds counts[s] entries
__global__ void __launch_bounds__(256) tree_apply_edges_kernel(const u64* __restrict__ queue, u64 count, u32 sigma, u64* __restrict__ nodes) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (u64)gridDim.x * blockDim.x)
        nodes[queue[3 * e] * (u64)(sigma + 1) + queue[3 * e + 1]] = queue[3 * e + 2];
}

template <typename T>
__global__ void __launch_bounds__(32) array_min_kernel(const T* __restrict__ a, u64 n, T* __restrict__ out) {
    T m = n ? a[0] : (T)0;
    for (u64 j = threadIdx.x; j < n; j += 32) m = a[j] < m ? a[j] : m;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const T o = __shfl_xor_sync(0xffffffffu, m, d);
        m = o < m ? o : m;
    }
    if (threadIdx.x == 0) *out = m;
}

// ---- host launchers of the tile kernels (dynamic shared memory above 48 KB needs the attribute once per device); the
// list kernel runs on a full grid: `sms` multiprocessors x 8 CTAs
template <typename T, class S>
void launch_ansv_tile(const S& sr, const T* vals, u64 g0, u64 m, int left_mode, int right_mode, u64 nonsv, u64* left, u64* right, AnsvList L, int sms,
                      cudaStream_t st) {
    auto kern = ansv_tile_kernel<T, S>;
    static bool seen[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (!seen[d & 63]) {
        seen[d & 63] = true;
        PSAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AnsvTile<T>::SMEM));
    }
    PSAC_CUDA(cudaMemsetAsync(L.count, 0, sizeof(u64), st));
    kern<<<(unsigned)((m + AnsvTile<T>::TILE - 1) / AnsvTile<T>::TILE), 256, AnsvTile<T>::SMEM, st>>>(sr, vals, g0, m, left_mode, right_mode, nonsv, left, right, L);
    ansv_list_kernel<T, S><<<sms * 8, 256, 0, st>>>(sr, g0, left_mode, right_mode, nonsv, left, right, L);
    PSAC_CUDA(cudaGetLastError());
}
// shared memory the staged rows may take next to the sparse table (the static s_feq array is on top of both)
constexpr size_t TREE_STAGE_BUDGET = 200 * 1024;
template <typename V>
static inline bool tree_rows_staged(u32 sigma) {
    return AnsvTile<V>::SMEM + (size_t)(sigma + 1) * AnsvTile<V>::TILE * sizeof(u64) <= TREE_STAGE_BUDGET && !getenv("PSACB200_NO_TREE_STAGE");
}
// does the tile kernel write every row of the table itself (no clearing pass needed)?
template <typename IdxT>
static inline bool tree_tile_writes_all_rows(u64 n, u32 sigma) {
    return (sizeof(IdxT) == 8 && n <= (1ull << 32)) ? tree_rows_staged<u32>(sigma) : tree_rows_staged<IdxT>(sigma);
}

template <typename IdxT, typename V, class S>
void launch_tree_tile_v(const TreeFusedArgs<IdxT>& A, const S& sr, AnsvList L, int sms, cudaStream_t st) {
    const bool stage = tree_rows_staged<V>(A.sigma);
    auto kern = suffix_tree_tile_kernel<IdxT, V, S, false>;
    auto kern_stage = suffix_tree_tile_kernel<IdxT, V, S, true>;
    static bool seen[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (!seen[d & 63]) {
        seen[d & 63] = true;
        PSAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AnsvTile<V>::SMEM));
        PSAC_CUDA(cudaFuncSetAttribute(kern_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TREE_STAGE_BUDGET));
    }
    PSAC_CUDA(cudaMemsetAsync(L.count, 0, sizeof(u64), st));
    const unsigned tiles = (unsigned)((A.m + AnsvTile<V>::TILE - 1) / AnsvTile<V>::TILE);
    // one CTA per SM (the shared memory is full): 1024 threads -- measured at 2^29: 512 threads 40.6 ms, 1024 threads 35.0 ms
    static const int stage_threads = getenv("PSACB200_TREE_THREADS") ? atoi(getenv("PSACB200_TREE_THREADS")) : 1024;
    if (stage)
        kern_stage<<<tiles, stage_threads, AnsvTile<V>::SMEM + (size_t)(A.sigma + 1) * AnsvTile<V>::TILE * sizeof(u64), st>>>(A, sr, L);
    else
        kern<<<tiles, 256, AnsvTile<V>::SMEM, st>>>(A, sr, L);
    tree_list_kernel<IdxT, S><<<sms * 8, 256, 0, st>>>(A, sr, L);
    PSAC_CUDA(cudaGetLastError());
}
template <typename IdxT, class S>
void launch_tree_tile(const TreeFusedArgs<IdxT>& A, const S& sr, AnsvList L, int sms, cudaStream_t st) {
    if (sizeof(IdxT) == 8 && A.n <= (1ull << 32))
        launch_tree_tile_v<IdxT, u32, S>(A, sr, L, sms, st);  // LCP values are below n: 32 bits in shared memory
    else
        launch_tree_tile_v<IdxT, IdxT, S>(A, sr, L, sms, st);
}

}  // namespace psacb200
