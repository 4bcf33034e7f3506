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

constexpr int ANSV_FAN = 32;
constexpr int ANSV_MAX_LEVELS = 9;
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

// Nearest position j on the LEFT (DIR = -1: j < i) or RIGHT (DIR = +1: j > i) of i with v[j] < x (STRICT) or <= x.
// ANSV_NONE if there is none.
template <typename T, int DIR, bool STRICT>
__device__ u64 ansv_search(const MinTree<T>& t, u64 i, T x) {
    int lv = 0;
    u64 p = i;  // current position at level lv; its block of 32 is [p & ~31, p | 31]
    // ---- climb: scan the rest of the current block away from p, then move to the parent
    while (true) {
        const u64 n = t.size[lv];
        const T* a = t.level[lv];
        u64 found = ANSV_NONE;
        if (DIR < 0) {
            const u64 lo = p & ~(u64)(ANSV_FAN - 1);
            for (u64 j = p; j > lo;) {
                --j;
                if (ansv_hit<T, STRICT>(a[j], x)) {
                    found = j;
                    break;
                }
            }
        } else {
            u64 hi = (p | (u64)(ANSV_FAN - 1)) + 1;
            hi = hi < n ? hi : n;
            for (u64 j = p + 1; j < hi; ++j) {
                if (ansv_hit<T, STRICT>(a[j], x)) {
                    found = j;
                    break;
                }
            }
        }
        if (found != ANSV_NONE) {
            p = found;
            break;
        }
        if (lv + 1 >= t.levels) return ANSV_NONE;
        p >>= 5;
        ++lv;
    }
    // ---- descend: inside block p of level lv take the child nearest to i that qualifies
    while (lv > 0) {
        --lv;
        const u64 n = t.size[lv];
        const T* a = t.level[lv];
        const u64 lo = p * ANSV_FAN;
        u64 hi = lo + ANSV_FAN;
        hi = hi < n ? hi : n;
        if (DIR < 0) {
            u64 j = hi;
            while (j > lo) {
                --j;
                if (ansv_hit<T, STRICT>(a[j], x)) break;
            }
            p = j;
        } else {
            u64 j = lo;
            while (j < hi && !ansv_hit<T, STRICT>(a[j], x)) ++j;
            p = j;
        }
    }
    return p;
}

// inside block-minimum tree t, the position nearest to the entry side (DIR < 0: we come from the right, so the LARGEST
// qualifying index; DIR > 0: the smallest) -- the caller knows that the minimum of the whole array qualifies
template <typename T, int DIR, bool STRICT>
__device__ u64 ansv_descend_top(const MinTree<T>& t, T x) {
    int lv = t.levels - 1;
    u64 p = 0;
    {
        const T* a = t.level[lv];
        const u64 n = t.size[lv];
        if (DIR < 0) {
            u64 j = n;
            while (j > 0) {
                --j;
                if (ansv_hit<T, STRICT>(a[j], x)) break;
            }
            p = j;
        } else {
            u64 j = 0;
            while (j + 1 < n && !ansv_hit<T, STRICT>(a[j], x)) ++j;
            p = j;
        }
    }
    while (lv > 0) {
        --lv;
        const u64 n = t.size[lv];
        const T* a = t.level[lv];
        const u64 lo = p * ANSV_FAN;
        u64 hi = lo + ANSV_FAN;
        hi = hi < n ? hi : n;
        if (DIR < 0) {
            u64 j = hi;
            while (j > lo) {
                --j;
                if (ansv_hit<T, STRICT>(a[j], x)) break;
            }
            p = j;
        } else {
            u64 j = lo;
            while (j < hi && !ansv_hit<T, STRICT>(a[j], x)) ++j;
            p = j;
        }
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

template <typename IdxT, class S>
__global__ void __launch_bounds__(256) suffix_tree_fused_kernel(TreeFusedArgs<IdxT> A, S sr) {
    const u64 n = A.n;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < A.m; i += (u64)gridDim.x * blockDim.x) {
        const u64 gi = A.g0 + i;
        const u64 lcp_i = A.lcp[i];
        const u64 sa_i = A.sa[i];
        // ---- leaf n + gi (suffix_tree.hpp:88-141)
        u64 parent, lcp_val;
        u64 lnsv = ANSV_NONE, left_val = 0;
        const bool need_left = gi > 0;
        if (need_left) {
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
        if (gi == 0 || lcp_i == 0) continue;
        const u64 rnsv = ansv_one<IdxT, +1>(sr, gi, 0);
        if (rnsv == ANSV_NONE) {
            if (left_val == lcp_i) continue;  // duplicate of the node further left
            tree_emit_dist<IdxT>(A, lnsv, gi, sa_i, left_val);
        } else {
            const u64 right_val = sr.value(rnsv);
            if (left_val >= right_val) {
                if (left_val == lcp_i) continue;
                tree_emit_dist<IdxT>(A, lnsv, gi, sa_i, left_val);
            } else {
                tree_emit_dist<IdxT>(A, rnsv, gi, sa_i, right_val);
            }
        }
    }
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

}  // namespace psacb200
