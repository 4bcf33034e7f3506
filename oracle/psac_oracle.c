/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of patflick/psac's SA/LCP
 * construction path, in plain C.
 *
 * This file is the parity oracle for the CUDA engine.  It is never linked
 * into, imported by or executed from the product (psac_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it.  Parity is PINNED: tests/test_oracle.py checks every function
 * here against the reference's golden vectors and
 * against the unmodified reference compiled at np=1 (oracle/_ref, see
 * oracle/ref_driver.cpp), with fixtures committed under tests/golden/.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference tree).  All arrays are uint64_t regardless of the
 * reference's index_t; `index_bits` (32|64) only enters where the reference's
 * arithmetic depends on sizeof(index_t) (k-mer width).  np=1 semantics: the
 * halo exchanges of the reference (left/right shifts) degenerate to "0 past
 * the end" / "rank 0 has no left neighbour".
 *
 * Also restated: construct_ss, the generalized suffix array of a string set (section 8 f2).
 *
 * Not restated: construct_msgs (bucket chasing, suffix_array.hpp:1032-1285).
 * It is a schedule optimisation -- final SA/ISA/LCP do not depend on it
 * (0-padded suffixes are strictly totally ordered) -- and oracle/_ref runs the
 * real one; tests compare this file's output with _ref for fast=true inputs.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;

/* ---------------------------------------------------------------- alphabet */

/* alphabet<char>::init_mapping_table / init_sizes  (include/alphabet.hpp:147-164).
 * lut[c] = 1 + rank of c among used byte values, stored in an 8-bit table:
 * with all 256 byte values present 0xFF gets 256 -> wraps to 0. */
void oracle_alphabet(const uint8_t* text, size_t n, uint8_t lut[256], unsigned* sigma, unsigned* bits_per_char) {
    u64 hist[256];
    memset(hist, 0, sizeof hist);
    for (size_t i = 0; i < n; ++i) hist[text[i]]++; /* alphabet.hpp:48-59 */
    uint16_t mapped = 1;
    unsigned s = 0;
    for (int c = 0; c < 256; ++c) {
        lut[c] = 0;
        if (hist[c]) {
            lut[c] = (uint8_t)mapped; /* uchar_type truncation, alphabet.hpp:136,160 */
            ++mapped;
            ++s;
        }
    }
    *sigma = s;
    /* ceillog2(sigma+1), bitops.hpp:147-152 */
    unsigned l = 0;
    while ((1u << l) < s + 1) ++l;
    *bits_per_char = l;
}

/* alphabet::chars_per_word + get_optimal_k at p ranks (alphabet.hpp:254-262, kmer.hpp:25-40);
 * index types are unsigned on this path, so bits_per_word = index_bits. */
unsigned oracle_optimal_k(unsigned bits_per_char, unsigned index_bits, size_t min_local_size, int p, unsigned k) {
    unsigned max_k = index_bits / bits_per_char;
    if (k == 0 || k > max_k) k = max_k;
    if (k >= min_local_size) {
        k = (unsigned)min_local_size;
        if (p == 1 && k > 1) k--;
    }
    return k;
}

/* ------------------------------------------------------------------ k-mers */

/* kmer_generation / for_each_kmer (include/kmer.hpp:86-127, 179-201) at np=1.
 * B[i] = sum_j code(T[i+j]) << l*(k-1-j), zero past the end; k == 1 copies the RAW chars
 * (sign-extended `char`, kmer.hpp:196-199). */
void oracle_kmer_generation(const uint8_t* text, size_t n, const uint8_t lut[256], unsigned l, unsigned k,
                            unsigned index_bits, u64* out) {
    u64 word_mask = index_bits == 64 ? ~(u64)0 : (((u64)1 << index_bits) - 1);
    if (k == 1) {
        for (size_t i = 0; i < n; ++i) out[i] = (u64)(int64_t)(int8_t)text[i] & word_mask;
        return;
    }
    u64 kmer_mask = (l * k >= 64) ? ~(u64)0 : (((u64)1 << (l * k)) - 1);
    kmer_mask &= word_mask;
    if (kmer_mask == 0) kmer_mask = word_mask;
    size_t pos = 0, o = 0;
    u64 kmer = 0;
    size_t pre = (size_t)(k - 1) < n ? (size_t)(k - 1) : n;
    for (size_t i = 0; i < pre; ++i) { kmer = ((kmer << l) | lut[text[pos++]]) & word_mask; }
    if (n < (size_t)(k - 1)) kmer = (kmer << (l * (k - 1 - n))) & word_mask;
    while (pos < n) {
        kmer = ((kmer << l) | lut[text[pos++]]) & kmer_mask;
        out[o++] = kmer;
    }
    for (size_t i = 0; i < pre; ++i) {
        kmer = (kmer << l) & kmer_mask;
        out[o++] = kmer;
    }
}

/* shift_vector at np=1 (include/shifting.hpp:32-122): B2[i] = B[i+h], 0 past the end */
void oracle_shift(const u64* b, size_t n, size_t h, u64* out) {
    for (size_t i = 0; i < n; ++i) out[i] = (i + h < n) ? b[i + h] : 0;
}

/* ------------------------------------------------------------ tuple sorting */

typedef struct { u64 v1, v2, idx; } tuple3;

static int cmp_tuple3(const void* a, const void* b) {
    const tuple3* x = (const tuple3*)a; const tuple3* y = (const tuple3*)b;
    if (x->v1 != y->v1) return x->v1 < y->v1 ? -1 : 1;
    if (x->v2 != y->v2) return x->v2 < y->v2 ? -1 : 1;
    /* the reference's comparator stops here (idxsort.hpp:47-49): order inside equal (v1,v2) is
     * implementation-defined; ties are broken by idx so the oracle is deterministic */
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

/* idxsort_vectors at np=1 (include/idxsort.hpp:22-83): sort (v1,v2,idx=i) by (v1,v2); v1,v2 are
 * permuted in place, idx (= SA) is returned. */
int oracle_idxsort(u64* v1, u64* v2, size_t n, u64* idx_out) {
    tuple3* t = (tuple3*)malloc((n ? n : 1) * sizeof(tuple3));
    if (!t) return -1;
    for (size_t i = 0; i < n; ++i) { t[i].v1 = v1[i]; t[i].v2 = v2[i]; t[i].idx = i; }
    qsort(t, n, sizeof(tuple3), cmp_tuple3);
    for (size_t i = 0; i < n; ++i) { v1[i] = t[i].v1; v2[i] = t[i].v2; idx_out[i] = t[i].idx; }
    free(t);
    return 0;
}

/* ---------------------------------------------------------------- rebucket */

/* rebucket + global_fill_where_zero at np=1 (include/bucketing.hpp:57-123, 21-53).
 * v1[i] <- 1-based index of the first element of i's (v1,v2)-bucket; returns the number of
 * buckets with more than one element and the number of elements in them. */
void oracle_rebucket(u64* v1, const u64* v2, size_t n, u64* unfinished_buckets, u64* unfinished_elements) {
    if (n == 0) { *unfinished_buckets = *unfinished_elements = 0; return; }
    int next_diff = 1; /* rank 0: firstDiff = true */
    for (size_t i = 0; i + 1 < n; ++i) {
        int set_one = next_diff;
        next_diff = !(v1[i] == v1[i + 1] && v2[i] == v2[i + 1]);
        v1[i] = set_one ? (u64)i + 1 : 0;
    }
    v1[n - 1] = next_diff ? (u64)(n - 1) + 1 : 0;
    u64 ub = 0, ue = 0;
    for (size_t i = 1; i < n; ++i) {
        if (v1[i - 1] > 0 && v1[i] == 0) { ++ub; ++ue; }
        if (v1[i] == 0) ++ue;
    }
    *unfinished_buckets = ub;
    *unfinished_elements = ue;
    u64 pre_max = 0;
    for (size_t i = 0; i < n; ++i) {
        if (v1[i] == 0) v1[i] = pre_max; else pre_max = v1[i];
    }
}

/* bulk_permute_inplace at np=1 (include/bulk_permute.hpp:14-73): out[idx[i]] = vec[i] */
void oracle_bulk_permute(const u64* vec, const u64* idx, size_t n, u64* out) {
    for (size_t i = 0; i < n; ++i) out[idx[i]] = vec[i];
}

/* --------------------------------------------------------------------- LCP */

static unsigned clz64(u64 x) { return x ? (unsigned)__builtin_clzll(x) : 64; }

/* lcp_bitwise (include/bitops.hpp:169-183); word_bits = 8*sizeof(T) */
unsigned oracle_lcp_bitwise(u64 x, u64 y, unsigned k, unsigned bits_per_char, unsigned word_bits) {
    if (x == y) return k;
    u64 z = x ^ y;
    unsigned lz = clz64(z) - (64 - word_bits);
    unsigned kmer_lz = lz - (word_bits - k * bits_per_char);
    return kmer_lz / bits_per_char;
}

/* initial_kmer_lcp at np=1 (include/suffix_array.hpp:1353-1396, _CONSTRUCT_LC=false branch).
 * b1,b2 = sorted k-mer pairs; LCP[i] = n ("unset") inside a bucket. */
void oracle_initial_kmer_lcp(const u64* b1, const u64* b2, size_t n, unsigned k, unsigned l, unsigned word_bits, u64* lcp) {
    for (size_t i = 0; i < n; ++i) lcp[i] = (u64)n;
    if (n) lcp[0] = 0;
    for (size_t i = 1; i < n; ++i) {
        if (b1[i - 1] != b1[i] || b2[i - 1] != b2[i]) {
            unsigned v = oracle_lcp_bitwise(b1[i - 1], b1[i], k, l, word_bits);
            if (v == k) v += oracle_lcp_bitwise(b2[i - 1], b2[i], k, l, word_bits);
            lcp[i] = v;
        }
    }
}

/* min over lcp[lo, hi) -- stands in for bulk_rmq_v2 (include/par_rmq.hpp:199-332), whose answer is
 * the range minimum; an iterative segment tree is enough for an oracle. */
typedef struct { size_t n; u64* t; } segtree;
static int seg_build(segtree* s, const u64* a, size_t n) {
    s->n = n; s->t = (u64*)malloc(2 * (n ? n : 1) * sizeof(u64));
    if (!s->t) return -1;
    for (size_t i = 0; i < n; ++i) s->t[n + i] = a[i];
    for (size_t i = n; i-- > 1;) s->t[i] = s->t[2 * i] < s->t[2 * i + 1] ? s->t[2 * i] : s->t[2 * i + 1];
    return 0;
}
static u64 seg_min(const segtree* s, size_t lo, size_t hi) {
    u64 r = ~(u64)0;
    for (lo += s->n, hi += s->n; lo < hi; lo >>= 1, hi >>= 1) {
        if (lo & 1) { if (s->t[lo] < r) r = s->t[lo]; ++lo; }
        if (hi & 1) { --hi; if (s->t[hi] < r) r = s->t[hi]; }
    }
    return r;
}

/* resolve_next_lcp at np=1 (include/suffix_array.hpp:1444-1508): for each NEW boundary
 * (b1 equal, b2 differ) LCP[i] = dist + min LCP[min(b2)..max(b2)); if a b2 is 0, LCP[i] = dist
 * when still unset.  All queries see the LCP of the previous round. */
int oracle_resolve_next_lcp(const u64* b1, const u64* b2, size_t n, u64 dist, u64* lcp) {
    segtree st;
    if (seg_build(&st, lcp, n)) return -1;
    for (size_t i = 1; i < n; ++i) {
        if (b1[i - 1] != b1[i]) continue;
        u64 l2 = b2[i - 1], r2 = b2[i];
        if (l2 == 0 || r2 == 0) {
            if (lcp[i] == (u64)n) lcp[i] = dist;
        } else if (l2 != r2) {
            u64 lo = l2 < r2 ? l2 : r2, hi = l2 < r2 ? r2 : l2;
            lcp[i] = dist + seg_min(&st, (size_t)lo, (size_t)hi); /* 0-based [lo, hi) on 1-based ids */
        }
    }
    free(st.t);
    return 0;
}

/* -------------------------------------------------------- construct (a1-a11) */

/* suffix_array::construct(begin,end,fast_resolval,k) at np=1, without the bucket-chasing
 * shortcut (include/suffix_array.hpp:365-466, 469-486).  sa/isa/lcp receive n entries each;
 * lcp may be NULL.  *rounds_out (optional) = number of doubling rounds executed. */
int oracle_construct(const uint8_t* text, size_t n, unsigned index_bits, unsigned k_in, u64* sa, u64* isa, u64* lcp,
                     unsigned* rounds_out) {
    if (n == 0) return 0;
    uint8_t lut[256];
    unsigned sigma, l;
    oracle_alphabet(text, n, lut, &sigma, &l);
    unsigned k = oracle_optimal_k(l, index_bits, n, 1, k_in); /* suffix_array.hpp:478-479 */
    u64* b = (u64*)malloc(n * sizeof(u64));
    u64* b2 = (u64*)malloc(n * sizeof(u64));
    if (!b || !b2) { free(b); free(b2); return -1; }
    oracle_kmer_generation(text, n, lut, l, k, index_bits, b); /* :370 */
    unsigned rounds = 0;
    u64 ub = 1, ue = n;
    int have_sa = 0;
    for (size_t h = k; h < n; h <<= 1) { /* :381 */
        oracle_shift(b, n, h, b2);                  /* :387 */
        if (oracle_idxsort(b, b2, n, sa)) { free(b); free(b2); return -1; } /* :394 */
        have_sa = 1;
        if (lcp) {                                  /* :401-409 */
            if (h == k) oracle_initial_kmer_lcp(b, b2, n, k, l, index_bits, lcp);
            else if (oracle_resolve_next_lcp(b, b2, n, (u64)h, lcp)) { free(b); free(b2); return -1; }
        }
        oracle_rebucket(b, b2, n, &ub, &ue);       /* :414 */
        oracle_bulk_permute(b, sa, n, b2);          /* :433-442 (SA is kept: it is final if this is the last round) */
        memcpy(b, b2, n * sizeof(u64));
        ++rounds;
        if (ub == 0) break;                         /* :448 */
    }
    if (!have_sa) { /* n == 1 (loop body never runs): the reference leaves SA empty; define SA = {0} */
        sa[0] = 0; b[0] = 1;
        if (lcp) lcp[0] = 0;
    }
    for (size_t i = 0; i < n; ++i) isa[i] = b[i] - 1; /* :460-464 */
    if (rounds_out) *rounds_out = rounds;
    free(b); free(b2);
    return ub == 0 ? 0 : 1; /* 1: loop ended with unfinished buckets (only with the sigma=256 code-0 quirk) */
}

/* ------------------------------------------- construct_ss: generalized SA (f2) */

/* simple_dstringset::parse at np=1 (include/stringset.hpp:42-80): maximal runs of non-separator characters are the
 * strings; str_end[i] (i = index in the concatenation WITHOUT separators) = exclusive end of i's string in that
 * concatenation; cat[] = the concatenation.  Returns sum_sizes. */
size_t oracle_parse_stringset(const uint8_t* flat, size_t len, uint8_t sep, uint8_t* cat, u64* str_end) {
    size_t m = 0, i = 0;
    while (i < len) {
        while (i < len && flat[i] == sep) ++i;
        size_t b = m;
        while (i < len && flat[i] != sep) cat[m++] = flat[i++];
        for (size_t j = b; j < m; ++j) str_end[j] = (u64)m;
    }
    return m;
}

/* kmer_gen_stringset at np=1 (include/kmer.hpp:269-355): k-mers never read past the end of their string (0 fill) */
void oracle_kmer_gen_stringset(const uint8_t* cat, const u64* str_end, size_t n, const uint8_t lut[256], unsigned l, unsigned k, u64* out) {
    for (size_t i = 0; i < n; ++i) {
        u64 v = 0;
        for (unsigned j = 0; j < k; ++j) {
            v <<= l;
            if (i + j < str_end[i]) v |= lut[cat[i + j]];
        }
        out[i] = v;
    }
}

/* shift_buckets_ds at np=1 (include/shifting.hpp:374-418): B2[i] = B[i+h] inside i's string, else 0 */
void oracle_shift_buckets_ds(const u64* b, const u64* str_end, size_t n, size_t h, u64* out) {
    for (size_t i = 0; i < n; ++i) out[i] = (i + h < str_end[i]) ? b[i + h] : 0;
}

static unsigned ctz64(u64 x) { return x ? (unsigned)__builtin_ctzll(x) : 64; }

/* initial_kmer_lcp_gsa at np=1 (include/suffix_array.hpp:1404-1441): as initial_kmer_lcp, but equal k-mers that end in
 * 0 fill (a string ended inside them) have the length of the string's remainder as their LCP */
void oracle_initial_kmer_lcp_gsa(const u64* b1, const u64* b2, size_t n, unsigned k, unsigned l, unsigned word_bits, u64* lcp) {
    for (size_t i = 0; i < n; ++i) lcp[i] = (u64)n;
    if (n) lcp[0] = 0;
    for (size_t i = 1; i < n; ++i) {
        u64 l1 = b1[i - 1], l2 = b2[i - 1], r1 = b1[i], r2 = b2[i];
        if (l1 != r1) {
            lcp[i] = oracle_lcp_bitwise(l1, r1, k, l, word_bits);
        } else {
            unsigned v = k - ctz64(l1) / l;
            if (v == k) {
                if (l2 != r2) {
                    lcp[i] = v + oracle_lcp_bitwise(l2, r2, k, l, word_bits);
                } else {
                    if (l2 != 0) v += k - ctz64(l2) / l;
                    if (v < 2 * k) lcp[i] = v;
                }
            } else {
                lcp[i] = v;
            }
        }
    }
}

/* rebucket_gsa_kmers / rebucket_gsa at np=1 (include/bucketing.hpp:130-143 on top of rebucket :57-123): two neighbours
 * share a bucket only if their tuples are equal AND the string has not ended inside them (last character of the
 * second k-mer non-zero in round 1 -- mask = low l bits --, second rank non-zero later -- mask = all ones). */
void oracle_rebucket_gsa(u64* v1, const u64* v2, size_t n, u64 mask, u64* unfinished_buckets, u64* unfinished_elements) {
    if (n == 0) { *unfinished_buckets = *unfinished_elements = 0; return; }
    int next_diff = 1;
    for (size_t i = 0; i + 1 < n; ++i) {
        int set_one = next_diff;
        next_diff = !(v1[i] == v1[i + 1] && v2[i] == v2[i + 1] && (v2[i] & mask) != 0);
        v1[i] = set_one ? (u64)i + 1 : 0;
    }
    v1[n - 1] = next_diff ? (u64)(n - 1) + 1 : 0;
    u64 ub = 0, ue = 0;
    for (size_t i = 1; i < n; ++i) {
        if (v1[i - 1] > 0 && v1[i] == 0) { ++ub; ++ue; }
        if (v1[i] == 0) ++ue;
    }
    *unfinished_buckets = ub;
    *unfinished_elements = ue;
    u64 pre_max = 0;
    for (size_t i = 0; i < n; ++i) {
        if (v1[i] == 0) v1[i] = pre_max; else pre_max = v1[i];
    }
}

/* suffix_array::construct_ss at np=1 (include/suffix_array.hpp:269-363): generalized suffix / LCP array of the strings
 * of `flat` (separated by `sep`), alphabet = lut/l as given by the caller (alphabet<char>::from_string).  Positions index
 * the concatenation without separators; identical suffixes of different strings are ordered by position (the sort is
 * stable, idxsort_vectors<..., true> :297).  The reference switches to bucket chasing after its first round (:326-334,
 * construct_msgs_gsa); this restatement keeps doubling with the same tuple / tie / LCP rules (rebucket_bucket_gsa :915-919
 * and the LCP rule of rebucket_bucket :862-875 are those of rebucket_gsa and resolve_next_lcp), so the results are
 * the same.  sa/isa/lcp receive sum_sizes entries (lcp may be NULL); returns sum_sizes or < 0. */
long oracle_construct_ss(const uint8_t* flat, size_t len, uint8_t sep, const uint8_t lut[256], unsigned l, unsigned index_bits, u64* sa, u64* isa,
                         u64* lcp, unsigned* rounds_out) {
    uint8_t* cat = (uint8_t*)malloc(len ? len : 1);
    u64* str_end = (u64*)malloc((len ? len : 1) * sizeof(u64));
    if (!cat || !str_end) { free(cat); free(str_end); return -1; }
    const size_t n = oracle_parse_stringset(flat, len, sep, cat, str_end);
    if (n == 0) { free(cat); free(str_end); return 0; }
    unsigned k = oracle_optimal_k(l, index_bits, n, 1, 0); /* :274 */
    u64* b = (u64*)malloc(n * sizeof(u64));
    u64* b2 = (u64*)malloc(n * sizeof(u64));
    if (!b || !b2) { free(cat); free(str_end); free(b); free(b2); return -1; }
    oracle_kmer_gen_stringset(cat, str_end, n, lut, l, k, b); /* :284 */
    unsigned rounds = 0;
    u64 ub = 1, ue = n;
    int have_sa = 0;
    for (size_t h = k; n > 1 && h < 4 * n + 64; h <<= 1) { /* :292; the chasing loop continues until nothing is active */
        oracle_shift_buckets_ds(b, str_end, n, h, b2);            /* :296 */
        if (oracle_idxsort(b, b2, n, sa)) { ub = 2; break; }       /* :300, stable */
        have_sa = 1;
        if (lcp) {                                                /* :304-315 */
            if (h == k) oracle_initial_kmer_lcp_gsa(b, b2, n, k, l, index_bits, lcp);
            else if (oracle_resolve_next_lcp(b, b2, n, (u64)h, lcp)) { ub = 2; break; }
        }
        oracle_rebucket_gsa(b, b2, n, h == k ? (((u64)1 << l) - 1) : ~(u64)0, &ub, &ue); /* :310, :316 */
        oracle_bulk_permute(b, sa, n, b2);                         /* :325-333 */
        memcpy(b, b2, n * sizeof(u64));
        ++rounds;
        if (ub == 0) break;
    }
    if (!have_sa) { sa[0] = 0; b[0] = 1; if (lcp) lcp[0] = 0; ub = 0; }
    for (size_t i = 0; i < n; ++i) isa[i] = b[i] - 1; /* :358-362 */
    if (rounds_out) *rounds_out = rounds;
    free(cat); free(str_end); free(b); free(b2);
    return ub == 0 ? (long)n : -2;
}

/* -------------------------------------------------- construct_arr<L> (a13) */

#define ORACLE_MAX_L 4
typedef struct { u64 v[ORACLE_MAX_L + 1]; } tupleL; /* v[0] = idx, v[1..L] = ranks, suffix_array.hpp:527-532 */
static int g_cmp_L;
static int cmp_tupleL(const void* a, const void* b) {
    const tupleL* x = (const tupleL*)a; const tupleL* y = (const tupleL*)b;
    for (int j = 1; j <= g_cmp_L; ++j)
        if (x->v[j] != y->v[j]) return x->v[j] < y->v[j] ? -1 : 1;
    if (x->v[0] != y->v[0]) return x->v[0] < y->v[0] ? -1 : 1; /* deterministic tie-break, see cmp_tuple3 */
    return 0;
}

/* suffix_array::construct_arr<L> at np=1 without chasing (include/suffix_array.hpp:490-641;
 * multi_shift_inplace shifting.hpp:126-240; rebucket_arr bucketing.hpp:171-251): tuples of L ranks
 * B[i], B[i+h], ..., B[i+(L-1)h]; h *= L per round.  SA/ISA only. */
int oracle_construct_arr(const uint8_t* text, size_t n, unsigned index_bits, int L, u64* sa, u64* isa, unsigned* rounds_out) {
    if (n == 0) return 0;
    if (L < 2 || L > ORACLE_MAX_L) return -2;
    uint8_t lut[256];
    unsigned sigma, l;
    oracle_alphabet(text, n, lut, &sigma, &l);
    unsigned k = oracle_optimal_k(l, index_bits, n, 1, 0);
    u64* b = (u64*)malloc(n * sizeof(u64));
    tupleL* t = (tupleL*)malloc(n * sizeof(tupleL));
    if (!b || !t) { free(b); free(t); return -1; }
    oracle_kmer_generation(text, n, lut, l, k, index_bits, b);
    unsigned rounds = 0;
    u64 ub = 1;
    int have_sa = 0;
    for (size_t h = k; h < n; h *= (size_t)L) {
        for (size_t i = 0; i < n; ++i) {
            t[i].v[0] = i;
            for (int j = 0; j < L; ++j) {
                size_t p = i + (size_t)j * h;
                t[i].v[1 + j] = p < n ? b[p] : 0;
            }
        }
        g_cmp_L = L;
        qsort(t, n, sizeof(tupleL), cmp_tupleL);
        /* rebucket_arr: head flag where any of the L ranks differs from the predecessor */
        u64 cur = 0; ub = 0;
        size_t run = 0;
        for (size_t i = 0; i < n; ++i) {
            int head = (i == 0);
            if (!head) for (int j = 1; j <= L; ++j) if (t[i].v[j] != t[i - 1].v[j]) { head = 1; break; }
            if (head) { if (run > 1) ++ub; run = 0; cur = (u64)i + 1; }
            ++run;
            sa[i] = t[i].v[0];
            b[t[i].v[0]] = cur; /* bulk_permute: ISA order */
        }
        if (run > 1) ++ub;
        have_sa = 1;
        ++rounds;
        if (ub == 0) break;
    }
    if (!have_sa) { sa[0] = 0; b[0] = 1; }
    for (size_t i = 0; i < n; ++i) isa[i] = b[i] - 1;
    if (rounds_out) *rounds_out = rounds;
    free(b); free(t);
    return ub == 0 ? 0 : 1;
}

/* ------------------------------------------------------ sequential checkers */

/* lcp_from_sa (Kasai; include/lcp.hpp:46-77) */
void oracle_lcp_from_sa(const uint8_t* s, size_t n, const u64* sa, const u64* isa, u64* lcp) {
    if (n == 0) return;
    lcp[0] = 0;
    size_t h = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t k = h > 0 ? h - 1 : 0;
        if (isa[i] > 0) {
            size_t j = (size_t)sa[isa[i] - 1];
            while (i + k < n && j + k < n && s[i + k] == s[j + k]) k++;
        } else {
            k = 0; /* the reference's loop condition fails immediately when ISA[i]==0 and h<=1; LCP[0]=0 by definition */
        }
        lcp[isa[i]] = k;
        h = k;
    }
    lcp[0] = 0;
}

/* Independent SA oracle: plain comparison sort of all suffixes under the REFERENCE's order,
 * i.e. by encoded characters lut[c] with the end of text (code 0) smallest.  Not taken from
 * the reference -- it is the definition the reference's output must satisfy (README.md:85-101). */
static const uint8_t* g_ns_text; static size_t g_ns_n; static const uint8_t* g_ns_lut;
static int cmp_suffix(const void* a, const void* b) {
    size_t x = (size_t)*(const u64*)a, y = (size_t)*(const u64*)b;
    if (x == y) return 0;
    size_t n = g_ns_n;
    while (x < n && y < n) {
        uint8_t cx = g_ns_lut[g_ns_text[x]], cy = g_ns_lut[g_ns_text[y]];
        if (cx != cy) return cx < cy ? -1 : 1;
        ++x; ++y;
    }
    if (x >= n && y >= n) return 0;
    /* one suffix ended: it reads as code 0 from here on */
    if (x >= n) { while (y < n) { if (g_ns_lut[g_ns_text[y]] != 0) return -1; ++y; } return 0; }
    while (x < n) { if (g_ns_lut[g_ns_text[x]] != 0) return 1; ++x; }
    return 0;
}
void oracle_sa_naive(const uint8_t* text, size_t n, u64* sa) {
    uint8_t lut[256]; unsigned sigma, l;
    oracle_alphabet(text, n, lut, &sigma, &l);
    for (size_t i = 0; i < n; ++i) sa[i] = i;
    g_ns_text = text; g_ns_n = n; g_ns_lut = lut;
    qsort(sa, n, sizeof(u64), cmp_suffix);
}

/* Linear-time SA certificate (the conditions of d_check_sa, include/check_suffix_array.hpp:190-267):
 * SA is a permutation, ISA its inverse, and for every i>0 with a=SA[i-1], b=SA[i]:
 * code(T[a]) < code(T[b]), or equal and rank(a+1) < rank(b+1) (rank of the empty suffix = -1).
 * Returns 0 if valid, else 1 + index of the first violation class. */
int oracle_check_sa(const uint8_t* text, size_t n, const uint8_t lut[256], const u64* sa, const u64* isa) {
    for (size_t i = 0; i < n; ++i) {
        if (sa[i] >= n) return 1;
        if (isa[sa[i]] != (u64)i) return 2;
    }
    for (size_t i = 1; i < n; ++i) {
        size_t a = (size_t)sa[i - 1], b = (size_t)sa[i];
        uint8_t ca = lut[text[a]], cb = lut[text[b]];
        if (ca > cb) return 3;
        if (ca == cb) {
            int64_t ra = a + 1 < n ? (int64_t)isa[a + 1] : -1;
            int64_t rb = b + 1 < n ? (int64_t)isa[b + 1] : -1;
            if (!(ra < rb)) return 4;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------- ANSV */

/* ansv_sequential (include/ansv.hpp:47-65): nearest strictly smaller value to the left (left=1)
 * or right (left=0); `nonsv` where none exists. */
int oracle_ansv_sequential(const u64* in, size_t n, int left, u64 nonsv, u64* nsv) {
    size_t* q = (size_t*)malloc((n ? n : 1) * sizeof(size_t));
    if (!q) return -1;
    size_t top = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t idx = left ? n - 1 - i : i;
        while (top > 0 && in[idx] < in[q[top - 1]]) { nsv[q[top - 1]] = idx; --top; }
        q[top++] = idx;
    }
    for (size_t j = 0; j < top; ++j) nsv[q[j]] = nonsv;
    free(q);
    return 0;
}

/* ansv<T,left_type,right_type,global_indexing> at np=1 (include/ansv.hpp:2042-2045; the meaning of the
 * three modes is fixed by test/test_ansv.cpp:35-133 check_ansv, by its use in suffix_tree.hpp:62-141 and
 * by the fixtures generated from the unmodified reference, tests/golden/ansv137.npz):
 *   nearest_sm  (0): nearest j with in[j] <  in[i]
 *   nearest_eq  (1): nearest j with in[j] <= in[i]
 *   furthest_eq (2): walking away from i until the first strictly smaller element s: the furthest j
 *                    with in[j] == in[i] before s if there is one; otherwise the far end of s's run of
 *                    equal values (furthest j' beyond s with in[j'] == in[s] and nothing smaller than
 *                    in[s] in between; s itself if the run has one element).
 * `nonsv` where no such element exists.  O(n^2) worst case: small inputs only. */
int oracle_ansv(const u64* in, size_t n, int left_type, int right_type, u64 nonsv, u64* left, u64* right) {
    for (int dir = 0; dir < 2; ++dir) {
        int type = dir == 0 ? left_type : right_type;
        u64* out = dir == 0 ? left : right;
        for (size_t i = 0; i < n; ++i) {
            u64 res = nonsv;
            size_t limit = dir == 0 ? i : n - 1 - i;
            size_t t;
            int have_eq = 0, have_sm = 0;
            for (t = 1; t <= limit; ++t) { /* step t away from i */
                size_t j = dir == 0 ? i - t : i + t;
                if (in[j] < in[i]) { if (!have_eq) res = j; have_sm = 1; break; }
                if (in[j] == in[i]) {
                    if (type == 1) { res = j; break; }
                    if (type == 2) { res = j; have_eq = 1; }
                }
            }
            if (type == 2 && !have_eq && have_sm) {
                u64 m = in[res];
                for (++t; t <= limit; ++t) {
                    size_t j = dir == 0 ? i - t : i + t;
                    if (in[j] < m) break;
                    if (in[j] == m) res = j;
                }
            }
            out[i] = res;
        }
    }
    return 0;
}
