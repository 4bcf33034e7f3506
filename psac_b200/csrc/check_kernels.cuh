// psac-b200: device-side certificate of a suffix array / inverse suffix array / LCP array.
//
// Reference restated: d_check_sa (include/check_suffix_array.hpp:206-267; conditions at :190-194) and the LCP part of
// gl_check_correct -> check_lcp (:151-185):
//   1) SA is a permutation of {0..n-1}            } checked together: SA[i] < n and ISA[SA[i]] == i for every i
//   2) ISA is the inverse permutation of SA       } (n positions, ISA is a function on [0,n) => SA is injective)
//   3) S[SA[i-1]] <= S[SA[i]]                       (characters compared by their reference alphabet codes)
//   4) S[SA[i-1]] == S[SA[i]]  =>  ISA[SA[i-1]+1] < ISA[SA[i]+1]   (the empty suffix ranks below everything)
//   5) LCP[0] == 0 and LCP[i] == lcp(SA[i-1], SA[i]) by direct comparison on the packed text from offset 0
// The reference sorts copies of SA and bulk-permutes to pair every position with S[SA[i]] and ISA[SA[i]+1]; here every
// position simply gathers them: ISA is block-distributed over the ranks and read through peer-mapped memory
// (sharded.cuh PeerArena), the packed text is replicated.  O(n) random reads; independent of how the engine built the arrays.
#pragma once
#include "common.cuh"
#include "sa_kernels.cuh"

namespace psacb200 {

struct CheckArgs {
    const void* sa;         // local block of SA (IdxT)
    const void* lcp;        // local block of LCP or null
    u64 pos0;               // global SA position of local element 0
    u64 m;                  // local elements
    u64 n;                  // text length
    u64 halo_sa;            // SA[pos0 - 1] (used when pos0 > 0)
    const u64* stream;      // packed text (whole text)
    int lbits;
    int padded;             // sigma = 256 quirk: comparisons run over the zero-padded sequences (see stream_lcp)
    int p;                  // ISA blocks
    const void* isa_blk[16];  // block r of the ISA (text positions blk.start(r) ..), possibly peer memory
    BlkDiv div;
    unsigned long long* bad;  // [0] out of range, [1] inverse, [2] order, [3] lcp, [4] smallest failing SA position
};

template <typename IdxT>
__device__ __forceinline__ u64 check_isa_at(const CheckArgs& A, u64 g) {
    if (A.p == 1) return (u64)reinterpret_cast<const IdxT*>(A.isa_blk[0])[g];
    u64 local;
    const u32 r = A.div.owner(g, &local);
    return (u64)reinterpret_cast<const IdxT*>(A.isa_blk[r])[local];
}

template <typename IdxT>
__global__ void __launch_bounds__(256) check_sa_kernel(CheckArgs A) {
    const IdxT* sa = reinterpret_cast<const IdxT*>(A.sa);
    const IdxT* lcp = reinterpret_cast<const IdxT*>(A.lcp);
    unsigned long long nbad[4] = {0, 0, 0, 0};
    u64 first_bad = ~0ull;
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < A.m; q += (u64)gridDim.x * blockDim.x) {
        const u64 i = A.pos0 + q;
        const u64 s = (u64)sa[q];
        bool bad = false;
        if (s >= A.n) {
            nbad[0] += 1;
            first_bad = i < first_bad ? i : first_bad;
            continue;
        }
        if (check_isa_at<IdxT>(A, s) != i) {
            nbad[1] += 1;
            bad = true;
        }
        if (i == 0) {
            if (lcp != nullptr && (u64)lcp[0] != 0) {
                nbad[3] += 1;
                bad = true;
            }
        } else {
            const u64 a = q ? (u64)sa[q - 1] : A.halo_sa;
            if (a < A.n) {  // (an out-of-range predecessor is counted at its own position)
                const u32 ca = (u32)stream_extract(A.stream, a, A.lbits, A.lbits), cb = (u32)stream_extract(A.stream, s, A.lbits, A.lbits);
                bool ok = ca < cb;
                if (ca == cb) {
                    const u64 ra = a + 1 < A.n ? check_isa_at<IdxT>(A, a + 1) + 1 : 0;
                    const u64 rb = s + 1 < A.n ? check_isa_at<IdxT>(A, s + 1) + 1 : 0;
                    ok = ra < rb;
                }
                if (!ok) {
                    nbad[2] += 1;
                    bad = true;
                }
                if (lcp != nullptr) {
                    const u64 l = stream_lcp(A.stream, A.n, A.lbits, a, s, 0, A.padded != 0);
                    if ((u64)lcp[q] != l) {
                        nbad[3] += 1;
                        bad = true;
                    }
                }
            }
        }
        if (bad) first_bad = i < first_bad ? i : first_bad;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (nbad[c]) atomicAdd(&A.bad[c], nbad[c]);
    if (first_bad != ~0ull) atomicMin(&A.bad[4], (unsigned long long)first_bad);
}

// ------------------------------------------------------------------ left-branching characters (the Lc by-product)
// Reference: suffix_array<..., _CONSTRUCT_LC = true>::local_Lc (include/suffix_array.hpp:212, 1365-1383, 1485-1495;
// bulk_rmq_Lc, include/par_rmq.hpp:334-481), the array the DESA index consumes (include/desa.hpp:296-312, which recomputes it
// the same way): Lc[i] = S[SA[i-1] + LCP[i]] -- the character of the LEFT neighbour at the first mismatch -- and '\0' where
// that position lies past the end of the text or i = 0.  A pure function of (text, SA, LCP): one gather per position.
template <typename IdxT>
__global__ void __launch_bounds__(256) lc_kernel(const u64* __restrict__ stream, int lbits, const u8* __restrict__ inv, u64 n, const IdxT* __restrict__ sa,
                                                 const IdxT* __restrict__ lcp, u64 pos0, u64 m, u64 halo_sa, u8* __restrict__ out) {
    __shared__ u8 s_inv[256];
    s_inv[threadIdx.x] = inv[threadIdx.x];
    __syncthreads();
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (u64)gridDim.x * blockDim.x) {
        u8 c = 0;
        if (pos0 + q > 0) {
            const u64 g = (q ? (u64)sa[q - 1] : halo_sa) + (u64)lcp[q];
            if (g < n) c = s_inv[(u32)stream_extract(stream, g, lbits, lbits)];
        }
        out[q] = c;
    }
}

}  // namespace psacb200
