/*
 * psacb200 -- C ABI of the B200-native suffix-array / LCP construction engine.
 *
 * Drop-in boundary for ONE path of patflick/psac: suffix_array<char_t,index_t,_LCP>::construct() /
 * construct_arr<L>() (reference include/suffix_array.hpp:365-486, 490-641).  The reference has no FFI for this
 * path (it is a header-only class template); these are the entry points a C++/cgo/ctypes binding of that class
 * would bind, see INTEGRATION.md.  Plain pointers and sizes only; every function returns 0 on success or a
 * negative psacb200_status, and psacb200_last_error() gives the message (the C++ shim in
 * include/psacb200/suffix_array.hpp rethrows it as std::runtime_error like the reference, suffix_array.hpp:226).
 *
 * There is NO CPU fallback: every entry point fails with PSACB200_ERR_CUDA when no sm_100a device is usable.
 */
#ifndef PSACB200_H
#define PSACB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psacb200_engine psacb200_engine;

enum psacb200_status {
    PSACB200_OK = 0,
    PSACB200_ERR_ARG = -1,   /* bad argument (index width too small for n, null pointer, ...) */
    PSACB200_ERR_CUDA = -2,  /* CUDA runtime failure, including "no device" */
    PSACB200_ERR_OOM = -3,   /* device memory exhausted */
    PSACB200_ERR_INTERNAL = -4
};

/* flags of psacb200_construct* */
enum {
    PSACB200_LCP = 1u,          /* also build the LCP array (template parameter _CONSTRUCT_LCP, suffix_array.hpp:170) */
    PSACB200_FAST_RESOLVAL = 2u /* reference's fast_resolval argument (suffix_array.hpp:470); accepted for signature
                                   parity -- the engine always restricts later rounds to unresolved buckets */
};

/* Per-call statistics (the reference prints these as TIMER / "unfinished buckets" lines, suffix_array.hpp:52-63, 415). */
typedef struct psacb200_stats {
    uint64_t n;
    uint32_t sigma;           /* distinct characters */
    uint32_t bits_per_char;   /* reference's l = ceil(log2(sigma+1)), alphabet.hpp:153 */
    uint32_t pack_bits;       /* bits per character in the engine's packed text */
    uint32_t key_chars;       /* characters in the first sort key */
    uint32_t rounds;          /* sorting rounds executed (1 = the first sort resolved everything) */
    uint32_t sort_passes;     /* radix digit passes of the first sort */
    uint32_t internal_index_bytes;
    uint32_t reserved;        /* sharded construction: 1 = the SA->ISA exchange used peer stores over NVLink (fused kernel) */
    uint64_t unresolved_after_first; /* suffixes still sharing a bucket after the first sort */
    uint64_t device_bytes;    /* device memory held by the engine */
    float ms_total;           /* device time of the whole call (CUDA events) */
    float ms_h2d, ms_alphabet, ms_pack, ms_keygen /* 0: fused into digit pass 1 */, ms_hist, ms_sort, ms_resolve, ms_rounds, ms_output, ms_d2h;
    float ms_sort_pass_avg;   /* average duration of one radix digit pass of the first sort, pass 1 excluded */
    float ms_isa;             /* SA -> ISA permutation of the first round (partition pass + windowed scatter) */
    float ms_sort_pass1;      /* digit pass 1 of the first sort (keys read from the packed text) */
    float ms_scatter_avg;     /* average duration of the scatter kernel of one segmented digit pass (8-byte elements) */
    uint32_t sort_elt_bytes;  /* bytes of the element (carried key + suffix index) that one digit pass of the first sort moves */
    uint32_t sharded_scheme;  /* sharded construction: 2 = word exchange fused into digit pass 1 + distributed rounds, 1 = key-range
                                 selection + replicated rounds (fallback), 0 = not sharded / replicated on every GPU */
} psacb200_stats;

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
/* One engine = one GPU (cudaSetDevice(device)) + one stream + reusable device buffers.  Not thread-safe per engine. */
int psacb200_create(int device, psacb200_engine** out);
void psacb200_destroy(psacb200_engine* e);
const char* psacb200_last_error(void);
/* Number of CUDA kernel launches issued by this engine since creation (bench.py's gpu_launches). */
uint64_t psacb200_launch_count(const psacb200_engine* e);
int psacb200_get_stats(const psacb200_engine* e, psacb200_stats* out);
/* Fine-grained device timeline of the last construct call: "label=milliseconds;..." for consecutive marks on the engine's
 * stream (CUDA events).  Synchronises the stream. */
int psacb200_trace(psacb200_engine* e, char* buf, size_t buf_len);
/* Ranking mode of the radix digit passes: 0 = one shared-memory atomic per key (relies on the lanes of one ATOMS instruction
 * being applied in lane order, verified on the device at every psacb200_create), 1 = match.any ranking (documented warp
 * primitives only; selected automatically when that self-test fails, or with the environment variable PSACB200_SAFE_RANK=1). */
int psacb200_rank_mode(const psacb200_engine* e, uint64_t* selftest_mismatches);
/* The engine's cudaStream_t (as void*), so a caller can bracket calls with its own CUDA events on that stream. */
void* psacb200_stream(const psacb200_engine* e);
/* Pre-size the device buffers for texts up to n characters (optional; avoids cudaMalloc inside a timed call). */
int psacb200_reserve(psacb200_engine* e, size_t n, int index_bytes, unsigned flags);

/* ---- alphabet (reference alphabet<char>::from_sequence, include/alphabet.hpp:147-164, 213-218) --------------- */
/* lut[c] = 1 + rank of byte c among the bytes that occur, truncated to 8 bits exactly like the reference
 * (0xFF wraps to 0 when all 256 values occur).  text is a HOST pointer. */
int psacb200_alphabet(psacb200_engine* e, const uint8_t* text, size_t n, uint8_t lut[256], uint32_t* sigma, uint32_t* bits_per_char);

/* ---- construct (reference suffix_array::construct, include/suffix_array.hpp:469-486) ------------------------- */
/* HOST buffers in and out.  index_bytes = sizeof(index_t) (4 or 8).  sa_out / isa_out: n elements each, 0-based
 * (the reference's local_SA / local_B at p = 1); lcp_out: n elements, LCP[0] = 0, required iff flags & PSACB200_LCP.
 * isa_out may be NULL.  k = reference's k argument (0 = automatic).  Equivalent to construct_arr<L> for SA/ISA
 * (suffix_array.hpp:490-641 produces the same final arrays). */
int psacb200_construct(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, unsigned flags, unsigned k, void* sa_out,
                       void* isa_out, void* lcp_out);
/* Same with the alphabet supplied by the caller (reference overload suffix_array.hpp:365-366).  lut = 256 codes as
 * produced by psacb200_alphabet; only their ORDER matters. */
int psacb200_construct_alphabet(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, unsigned flags, unsigned k,
                                const uint8_t lut[256], void* sa_out, void* isa_out, void* lcp_out);
/* DEVICE buffers in and out (text already resident in HBM; outputs stay there).
 * Stream contract of every *_device and *_sharded entry point: the engine reads its inputs and writes its outputs on ITS OWN
 * stream (psacb200_stream) and returns after synchronising that stream; whatever produced the inputs on other streams must have
 * completed before the call.  Output buffers that are 16-byte aligned are used in place as the engine's result arrays (also when
 * the caller's index is 64 bits wide and the engine's 32: the narrow array lives in the lower half and is widened in place). */
int psacb200_construct_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, unsigned flags, unsigned k, void* d_sa,
                              void* d_isa, void* d_lcp);

/* ---- texts over wide characters (reference suffix_array<int, index_t, LCP>::construct with int_alphabet,
 *      include/alphabet.hpp:355-513; test/test_psac.cpp:277-304 "IntAlphabetMiss") ------------------------------------ */
/* text: n characters of char_bytes (2 or 4) bytes each, signed (char_signed != 0) or unsigned; the characters are ordered by
 * VALUE.  At most 255 distinct values may occur (else PSACB200_ERR_ARG): the text is reduced on the device to one byte per
 * character (1 + rank of its value) and the byte construction runs on that -- SA / ISA / LCP depend only on the order and
 * equality of the characters.  distinct_out (optional, room for 255) receives the values that occur in ascending order,
 * *n_distinct their number (min / max of the reference's int_alphabet = first / last).  Outputs as psacb200_construct.
 * (Not reproduced: with a 32-bit index_t and more than 16 bits per character the reference falls to k = 1, copies the raw,
 * sign-extended characters (include/kmer.hpp:196-199) and misorders negative values.) */
int psacb200_construct_wide(psacb200_engine* e, const void* text, size_t n, int char_bytes, int char_signed, int index_bytes, unsigned flags,
                            unsigned k, void* sa_out, void* isa_out, void* lcp_out, int64_t* distinct_out, uint32_t* n_distinct);
/* Same on DEVICE buffers (distinct_out / n_distinct stay host pointers). */
int psacb200_construct_wide_device(psacb200_engine* e, const void* d_text, size_t n, int char_bytes, int char_signed, int index_bytes,
                                   unsigned flags, unsigned k, void* d_sa, void* d_isa, void* d_lcp, int64_t* distinct_out, uint32_t* n_distinct);

/* ---- generalized suffix array of a string set (reference suffix_array::construct_ss, include/suffix_array.hpp:269-363,
 *      on a simple_dstringset, include/stringset.hpp:33-152; golden vectors test/test_gsa.cpp:97-98) ---------------- */
/* flat = the strings separated by the byte `sep` (runs of separators, leading and trailing ones are allowed and skipped, as
 * simple_dstringset::parse does).  The result indexes the concatenation of the strings WITHOUT separators: *n_out = number of
 * non-separator characters (the reference's ss.sum_sizes), and sa_out / isa_out / lcp_out receive *n_out entries each (the
 * caller provides room for len).  A suffix ends with its string; identical suffixes of different strings are ordered by
 * position and their LCP is their length.  lut: NULL = alphabet of the characters that occur (alphabet<char>::from_string of
 * all strings), or 256 codes whose ORDER defines the character order (codes of occurring characters must be non-zero).
 * One GPU; isa_out may be NULL. */
int psacb200_construct_ss(psacb200_engine* e, const uint8_t* flat, size_t len, uint8_t sep, int index_bytes, unsigned flags, const uint8_t* lut,
                          void* sa_out, void* isa_out, void* lcp_out, uint64_t* n_out);
/* Same on DEVICE buffers (stream contract as psacb200_construct_device). */
int psacb200_construct_ss_device(psacb200_engine* e, const uint8_t* d_flat, size_t len, uint8_t sep, int index_bytes, unsigned flags,
                                 const uint8_t* lut, void* d_sa, void* d_isa, void* d_lcp, uint64_t* n_out);

/* ---- building blocks on DEVICE pointers (used by the parity tests and by the sharded multi-GPU driver) ------ */
/* Stable LSD radix sort of n (key, value) pairs by key bits [begin_bit, end_bit); key_bytes in {4,8}, val_bytes in
 * {0,4,8}.  Sorted data is returned in keys/vals (keys_alt/vals_alt are scratch of the same size).
 * Replaces idxsort_vectors -> mxx::sort (include/idxsort.hpp:22-83). */
int psacb200_sort_pairs(psacb200_engine* e, void* d_keys, void* d_keys_alt, void* d_vals, void* d_vals_alt, size_t n, int key_bytes,
                        int val_bytes, int begin_bit, int end_bit);
/* Host-buffer wrapper of the above for tests. */
int psacb200_sort_pairs_host(psacb200_engine* e, void* keys, void* vals, size_t n, int key_bytes, int val_bytes, int begin_bit, int end_bit);

/* ---- sharded construction over the GPUs of one box: one process (rank) per GPU -------------------------------- */
/* The reference is SPMD over an MPI communicator (suffix_array(const mxx::comm&), include/suffix_array.hpp:174): every
 * rank calls construct() with its block of the text and ends up with its blocks of SA / ISA / LCP under mxx::blk_dist
 * (ext/mxx/include/mxx/partition.hpp:283-331).  Here a rank is a process driving one GPU; the exchange steps run over
 * NCCL on NVLink.  Rank 0 calls psacb200_comm_unique_id, the 128 bytes are handed to every rank by the caller's own
 * channel (torch.distributed / MPI / a file), then every rank calls psacb200_comm_init on its engine (collective). */
int psacb200_comm_unique_id(uint8_t id[128]);
int psacb200_comm_init(psacb200_engine* e, const uint8_t id[128], int rank, int world);
/* Collective.  d_text_local: this rank's block of the text (DEVICE), n_local = its size, which must be the blk_dist size
 * for (n_global, world, rank) -- otherwise PSACB200_ERR_ARG, like the reference's std::runtime_error (suffix_array.hpp:226).
 * Outputs (DEVICE, n_local elements of index_bytes each): this rank's blocks of SA, ISA (may be NULL) and LCP. */
int psacb200_construct_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, unsigned flags,
                               unsigned k, void* d_sa_local, void* d_isa_local, void* d_lcp_local);
/* Collective, optional: releases the peer-visible memory of the sharded construction in an ordered way (every rank closes its
 * mappings of the peers' memory before any rank frees its own).  Call on every rank before psacb200_destroy. */
int psacb200_comm_finalize(psacb200_engine* e);

/* ---- device-side certificate of a result (reference d_check_sa, include/check_suffix_array.hpp:190-194, 206-267, plus the LCP
 * part of gl_check_correct / check_lcp, :151-185) -------------------------------------------------------------------- */
typedef struct psacb200_check_report {
    uint64_t n;
    uint64_t bad_range;   /* positions i with SA[i] >= n */
    uint64_t bad_inverse; /* positions with ISA[SA[i]] != i          (conditions 1 + 2: SA is a permutation, ISA its inverse) */
    uint64_t bad_order;   /* positions violating S[SA[i-1]] < S[SA[i]] or (equal and ISA[SA[i-1]+1] < ISA[SA[i]+1])   (3 + 4) */
    uint64_t bad_lcp;     /* positions with LCP[i] != lcp(SA[i-1], SA[i]) (direct comparison on the text), LCP[0] != 0 */
    uint64_t first_bad;   /* smallest failing SA position, ~0 when the result is correct */
    float ms;             /* device time of the check */
    uint32_t checked_lcp;
} psacb200_check_report;
/* All pointers are DEVICE buffers; d_lcp may be NULL (SA / ISA only).  Returns PSACB200_OK when the check RAN; the
 * verdict is in the report (all four counters zero = correct). */
int psacb200_check_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, const void* d_sa, const void* d_isa, const void* d_lcp,
                          psacb200_check_report* report);
/* Same from HOST arrays (copies them to the device first). */
int psacb200_check(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, const void* sa, const void* isa, const void* lcp,
                   psacb200_check_report* report);
/* Collective over the ranks of psacb200_comm_init: every rank passes its blocks (DEVICE); the report is the same on all ranks. */
int psacb200_check_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, const void* d_sa_local,
                           const void* d_isa_local, const void* d_lcp_local, psacb200_check_report* report);

/* ---- left-branching characters (reference template parameter _CONSTRUCT_LC: local_Lc, include/suffix_array.hpp:212, 1365-1383,
 * 1485-1495; consumed by the DESA index, include/desa.hpp:296-312) ------------------------------------------------------ */
/* lc[i] = text[SA[i-1] + LCP[i]], '\0' where that position is past the end of the text, and for i = 0.  A by-product of SA + LCP:
 * one gather per position.  DEVICE arrays / HOST arrays / collective over the ranks (blocks of SA and LCP in, block of Lc out). */
int psacb200_lc_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, const void* d_sa, const void* d_lcp, uint8_t* d_lc);
int psacb200_lc(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, const void* sa, const void* lcp, uint8_t* lc);
int psacb200_lc_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, const void* d_sa_local,
                        const void* d_lcp_local, uint8_t* d_lc_local);

/* Host-side plans of the sharded construction (no GPU needed; used by the CPU tests): mxx::blk_dist, and the splitters
 * over a key-prefix histogram (rank r sorts the bins [first[r], first[r+1]); first has p+1 entries, count p). */
void psacb200_blk_dist(uint64_t n, int p, int r, uint64_t* start, uint64_t* size);
int psacb200_choose_splitters(const uint64_t* hist, size_t nbins, uint64_t n, int p, uint64_t* first, uint64_t* count);
/* Plan of the exchange that is fused into the first digit pass of the sharded sort: cnt[s * nb + d] = suffixes of text block s
 * whose top key digit is d.  Outputs: first[p+1] / cnt_key[p] (rank r sorts the digits [first[r], first[r+1])), owner[nb],
 * seg_dense / seg_pad [p][257] (dense and tile-padded start of every digit's segment in its owner's buffer), run_off[p][nb]
 * (where source s's run of digit d starts inside that buffer; sources are laid out p-1, 0, 1, .., p-2), balanced (0 = some
 * rank would get nothing or more than twice its share: the caller uses the fallback scheme). */
int psacb200_plan_word_exchange(const uint64_t* cnt, int p, int nb, uint64_t n, uint64_t pad_tile, uint64_t* first, uint64_t* cnt_key, int32_t* owner,
                                uint64_t* seg_dense, uint64_t* seg_pad, uint64_t* run_off, int* balanced);

/* ---- ANSV and suffix tree (reference include/ansv.hpp:2042-2051, include/suffix_tree.hpp:413-499) ------------------- */
/* All nearest smaller values of n HOST values (val_bytes 4 or 8): left[i] / right[i] = index of the match on that side,
 * or `nonsv` (global indexing at p = 1).  Modes: 0 nearest_sm, 1 nearest_eq, 2 furthest_eq (ansv.hpp:24-45). */
int psacb200_ansv(psacb200_engine* e, const void* vals, size_t n, int val_bytes, int left_type, int right_type, uint64_t nonsv, uint64_t* left,
                  uint64_t* right);
/* Suffix tree from SA + LCP (HOST arrays of index_bytes each) and the text, as the reference's child table
 * (construct_suffix_tree, suffix_tree.hpp:440-499): nodes[(sigma+1) * parent + code] = child, internal nodes are LCP
 * indices, leaves are n + SA position, 0 = empty.  nodes_len >= (sigma+1) * n (sigma from psacb200_alphabet). */
int psacb200_suffix_tree(psacb200_engine* e, const uint8_t* text, size_t n, int index_bytes, const void* sa, const void* lcp, uint64_t* nodes,
                         size_t nodes_len);

/* ---- several GPUs of one box behind ONE host call (the object a p = 1 caller of the reference class binds) -- */
/* One engine per GPU, driven by one host thread each inside the call; text, SA, ISA and LCP are sharded by block over the GPUs
 * internally (the same sharded construction as psacb200_construct_sharded; peer memory through plain peer access), the caller
 * passes and receives WHOLE HOST arrays exactly as with psacb200_construct.  dev_ids may be NULL (devices 0 .. n_gpus-1). */
typedef struct psacb200_multi psacb200_multi;
int psacb200_multi_create(int n_gpus, const int* dev_ids, psacb200_multi** out);
void psacb200_multi_destroy(psacb200_multi* m);
int psacb200_multi_gpus(const psacb200_multi* m);
psacb200_engine* psacb200_multi_engine(const psacb200_multi* m, int r); /* engine of GPU r (owned by the handle), e.g. for psacb200_alphabet */
int psacb200_multi_construct(psacb200_multi* m, const uint8_t* text, size_t n, int index_bytes, unsigned flags, unsigned k, void* sa_out, void* isa_out,
                             void* lcp_out);
int psacb200_multi_get_stats(const psacb200_multi* m, psacb200_stats* out); /* statistics of GPU 0's share of the last call */

/* ---- the same on DEVICE arrays, on one GPU or block-sharded over the ranks of psacb200_comm_init (BASELINE configs[4]) ------- */
/* ANSV of n DEVICE values; left / right: DEVICE arrays of n u64. */
int psacb200_ansv_device(psacb200_engine* e, const void* d_vals, size_t n, int val_bytes, int left_type, int right_type, uint64_t nonsv, uint64_t* d_left,
                         uint64_t* d_right);
/* Collective: the values are block-distributed like mxx::blk_dist; every rank gets the matches of ITS block as GLOBAL indices
 * (reference ansv<T, left, right, global_indexing>, include/ansv.hpp:2042-2051).  Elements whose match lies in another rank's block
 * (the reference's unmatched prefix / suffix minima, ansv.hpp:1362-1442) are resolved through peer memory. */
int psacb200_ansv_sharded(psacb200_engine* e, const void* d_vals_local, size_t n_local, size_t n_global, int val_bytes, int left_type, int right_type,
                          uint64_t nonsv, uint64_t* d_left_local, uint64_t* d_right_local);
/* Child table of the suffix tree from DEVICE SA + LCP (index_bytes each) and the text; the ANSV (furthest_eq left, nearest_sm
 * right, reference suffix_tree.hpp:62) is searched on the fly, not materialised.  d_nodes: DEVICE, (sigma + 1) * n u64 (the rows of
 * all LCP indices); *sigma_out receives sigma. */
int psacb200_suffix_tree_device(psacb200_engine* e, const uint8_t* d_text, size_t n, int index_bytes, const void* d_sa, const void* d_lcp, uint64_t* d_nodes,
                                size_t nodes_len, uint32_t* sigma_out);
/* Collective: every rank passes its blocks of text / SA / LCP and receives the rows of the LCP indices of ITS block,
 * (sigma + 1) * n_local u64 (reference construct_suffix_tree over a communicator, suffix_tree.hpp:440-499; edges whose parent row
 * lives on another rank travel through peer memory instead of the reference's all-to-all). */
int psacb200_suffix_tree_sharded(psacb200_engine* e, const uint8_t* d_text_local, size_t n_local, size_t n_global, int index_bytes, const void* d_sa_local,
                                 const void* d_lcp_local, uint64_t* d_nodes_local, size_t nodes_len, uint32_t* sigma_out);

#ifdef __cplusplus
}
#endif
#endif /* PSACB200_H */
