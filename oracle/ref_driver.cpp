/*
 * TEST INFRASTRUCTURE ONLY -- C entry points around the UNMODIFIED reference.
 *
 * Compiles patflick/psac's own headers where they lie under /root/reference
 * (nothing is copied into this repo) against oracle/mpi_shim/mpi.h and runs
 * them at np=1.  Output: oracle/_ref/libpsacref.so (git-ignored).  Only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
 *
 * Reference entry points wrapped here:
 *   suffix_array<char,index_t,LCP>::construct      include/suffix_array.hpp:469-486
 *   suffix_array<...>::construct_arr<L>            include/suffix_array.hpp:490-641
 *   alphabet<char>::from_sequence / encode          include/alphabet.hpp:213-218, 274-278
 *   kmer_generation                                 include/kmer.hpp:204-224
 *   lcp_bitwise                                     include/bitops.hpp:169-183
 *   lcp_from_sa (Kasai checker)                     include/lcp.hpp:46-77
 *   ansv / ansv_sequential                          include/ansv.hpp:47-65, 2042-2051
 *   construct_suffix_tree                           include/suffix_tree.hpp:413-499
 *   suffix_array<int,...>::construct (int_alphabet)  include/alphabet.hpp:355-513
 *   suffix_array<...>::construct_ss (generalized SA) include/suffix_array.hpp:269-363, stringset.hpp:33-152
 */
#include <mpi.h>

#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <mxx/env.hpp>
#include <mxx/comm.hpp>

#include <suffix_array.hpp>
#include <alphabet.hpp>
#include <kmer.hpp>
#include <bitops.hpp>
#include <lcp.hpp>
#include <ansv.hpp>
#include <suffix_tree.hpp>

namespace {

/* the reference logs to std::cerr unconditionally (suffix_array.hpp:49); mute it */
struct cerr_mute {
    std::streambuf* old;
    std::ostringstream sink;
    bool on;
    explicit cerr_mute(bool enable) : old(nullptr), on(enable) { if (on) old = std::cerr.rdbuf(sink.rdbuf()); }
    ~cerr_mute() { if (on) std::cerr.rdbuf(old); }
};

bool g_verbose = false;

template <typename index_t, bool LCP>
int run_construct(const char* text, size_t n, unsigned k, int fast, int arr_L, void* sa_out, void* isa_out, void* lcp_out) {
    cerr_mute mute(!g_verbose);
    mxx::comm c;
    suffix_array<char, index_t, LCP> sa(c);
    switch (arr_L) {
        case 0: sa.construct(text, text + n, fast != 0, k); break;
        case 2: sa.template construct_arr<2>(text, text + n, fast != 0); break;
        case 3: sa.template construct_arr<3>(text, text + n, fast != 0); break;
        case 4: sa.template construct_arr<4>(text, text + n, fast != 0); break;
        default: return -2;
    }
    if (sa.local_SA.size() != n || sa.local_B.size() != n) return -3;
    if (sa_out) std::memcpy(sa_out, sa.local_SA.data(), n * sizeof(index_t));
    if (isa_out) std::memcpy(isa_out, sa.local_B.data(), n * sizeof(index_t));
    if (LCP && lcp_out) {
        if (sa.local_LCP.size() != n) return -4;
        std::memcpy(lcp_out, sa.local_LCP.data(), n * sizeof(index_t));
    }
    return 0;
}

/* Generalized suffix array of a string set (suffix_array.hpp:269-363): `flat` holds the strings separated by `sep`
 * (simple_dstringset, stringset.hpp:33-152; runs of separators and leading / trailing ones are skipped); the alphabet is
 * alphabet<char>::from_string(alpha_chars) as in test/test_gsa.cpp:86.  Outputs have sum_sizes entries (characters that
 * are not separators); returns that count, or < 0. */
template <typename index_t>
long run_construct_ss(const char* flat, size_t len, char sep, const char* alpha_chars, size_t n_alpha, void* sa_out, void* isa_out, void* lcp_out,
                      size_t cap) {
    cerr_mute mute(!g_verbose);
    mxx::comm c;
    std::string f(flat, len);
    simple_dstringset ss(f.begin(), f.end(), c, sep);
    alphabet<char> a = alphabet<char>::from_string(std::string(alpha_chars, n_alpha), c);
    suffix_array<char, index_t, true> sa(c);
    sa.construct_ss(ss, a);
    const size_t m = sa.local_SA.size();
    if (m > cap) return -3;
    if (sa_out) std::memcpy(sa_out, sa.local_SA.data(), m * sizeof(index_t));
    if (isa_out && sa.local_B.size() == m) std::memcpy(isa_out, sa.local_B.data(), m * sizeof(index_t));
    if (lcp_out && sa.local_LCP.size() == m) std::memcpy(lcp_out, sa.local_LCP.data(), m * sizeof(index_t));
    return (long)m;
}


/* suffix_array<int, index_t, LCP>::construct at np=1: texts over wide characters go through int_alphabet
 * (include/alphabet.hpp:355-513; test/test_psac.cpp:277-304 "IntAlphabetMiss") */
template <typename index_t, bool LCP>
int run_construct_int(const int* text, size_t n, void* sa_out, void* isa_out, void* lcp_out) {
    cerr_mute mute(!g_verbose);
    mxx::comm c;
    suffix_array<int, index_t, LCP> sa(c);
    std::vector<int> t(text, text + n);
    sa.construct(t.begin(), t.end());
    if (sa.local_SA.size() != n || sa.local_B.size() != n) return -3;
    if (sa_out) std::memcpy(sa_out, sa.local_SA.data(), n * sizeof(index_t));
    if (isa_out) std::memcpy(isa_out, sa.local_B.data(), n * sizeof(index_t));
    if (LCP && lcp_out) std::memcpy(lcp_out, sa.local_LCP.data(), n * sizeof(index_t));
    return 0;
}

} // namespace

extern "C" {

void psacref_set_verbose(int v) { g_verbose = v != 0; }

/* suffix_array<char,index_t,LCP>::construct / construct_arr<L> at np=1.
 * index_bytes 4|8; want_lcp 0|1 (ignored for arr_L != 0: the reference builds no LCP there);
 * arr_L = 0 -> construct(begin,end,fast,k); 2..4 -> construct_arr<L>(begin,end,fast). */
int psacref_construct(const char* text, size_t n, int index_bytes, int want_lcp, unsigned k, int fast, int arr_L,
                      void* sa_out, void* isa_out, void* lcp_out) {
    try {
        if (index_bytes == 4) {
            return want_lcp && arr_L == 0 ? run_construct<uint32_t, true>(text, n, k, fast, arr_L, sa_out, isa_out, lcp_out)
                                          : run_construct<uint32_t, false>(text, n, k, fast, arr_L, sa_out, isa_out, lcp_out);
        } else if (index_bytes == 8) {
            return want_lcp && arr_L == 0 ? run_construct<uint64_t, true>(text, n, k, fast, arr_L, sa_out, isa_out, lcp_out)
                                          : run_construct<uint64_t, false>(text, n, k, fast, arr_L, sa_out, isa_out, lcp_out);
        }
        return -1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_construct: %s\n", e.what());
        return -10;
    }
}

/* alphabet<char>::from_sequence: lut[c] = encode(c) (uchar, wraps at 256), sigma, bits_per_char */
int psacref_alphabet(const char* text, size_t n, uint8_t* lut256, unsigned* sigma, unsigned* bits_per_char) {
    alphabet<char> a = alphabet<char>::from_sequence(text, text + n);
    for (int ch = 0; ch < 256; ++ch) lut256[ch] = a.encode((char)ch);
    *sigma = a.sigma();
    *bits_per_char = a.bits_per_char();
    return 0;
}

/* get_optimal_k at np=1 (kmer.hpp:25-40) */
unsigned psacref_optimal_k(const char* text, size_t n, int index_bytes, unsigned k) {
    mxx::comm c;
    alphabet<char> a = alphabet<char>::from_sequence(text, text + n);
    return index_bytes == 4 ? get_optimal_k<uint32_t>(a, n, c, k) : get_optimal_k<uint64_t>(a, n, c, k);
}

/* kmer_generation at np=1 (kmer.hpp:204-224), k as given (caller clamps) */
int psacref_kmer_generation(const char* text, size_t n, int index_bytes, unsigned k, void* out) {
    mxx::comm c;
    alphabet<char> a = alphabet<char>::from_sequence(text, text + n);
    if (index_bytes == 4) {
        std::vector<uint32_t> v = kmer_generation<uint32_t>(text, text + n, k, a, c);
        std::memcpy(out, v.data(), n * 4);
    } else {
        std::vector<uint64_t> v = kmer_generation<uint64_t>(text, text + n, k, a, c);
        std::memcpy(out, v.data(), n * 8);
    }
    return 0;
}

unsigned psacref_lcp_bitwise32(uint32_t x, uint32_t y, unsigned k, unsigned l) { return lcp_bitwise<uint32_t>(x, y, k, l); }
unsigned psacref_lcp_bitwise64(uint64_t x, uint64_t y, unsigned k, unsigned l) { return lcp_bitwise<uint64_t>(x, y, k, l); }

/* Kasai checker (lcp.hpp:46-77), 64-bit indices */
int psacref_lcp_from_sa(const char* text, size_t n, const uint64_t* sa, const uint64_t* isa, uint64_t* lcp) {
    std::string s(text, n);
    std::vector<uint64_t> SA(sa, sa + n), ISA(isa, isa + n), LCP;
    lcp_from_sa(s, SA, ISA, LCP);
    std::memcpy(lcp, LCP.data(), n * 8);
    return 0;
}

/* ansv_sequential (ansv.hpp:47-65) */
int psacref_ansv_sequential(const uint64_t* in, size_t n, int left, uint64_t nonsv, uint64_t* out) {
    std::vector<uint64_t> v(in, in + n);
    std::vector<size_t> r = ansv_sequential(v, left != 0, (size_t)nonsv);
    for (size_t i = 0; i < n; ++i) out[i] = r[i];
    return 0;
}

/* ansv<T,left_type,right_type,global_indexing> at np=1 (ansv.hpp:2042-2045); modes: 0 nearest_sm, 1 nearest_eq, 2 furthest_eq */
int psacref_ansv(const uint64_t* in, size_t n, int left_type, int right_type, uint64_t nonsv, uint64_t* left_out, uint64_t* right_out) {
    cerr_mute mute(!g_verbose);
    mxx::comm c;
    std::vector<uint64_t> v(in, in + n);
    std::vector<size_t> l, r;
    std::vector<std::pair<uint64_t, size_t>> lr;
#define PSACREF_ANSV_CASE(LT, RT) \
    if (left_type == LT && right_type == RT) { ansv<uint64_t, LT, RT, global_indexing>(v, l, r, lr, c, (size_t)nonsv); }
    PSACREF_ANSV_CASE(nearest_sm, nearest_sm) else PSACREF_ANSV_CASE(nearest_sm, nearest_eq) else PSACREF_ANSV_CASE(nearest_sm, furthest_eq)
    else PSACREF_ANSV_CASE(nearest_eq, nearest_sm) else PSACREF_ANSV_CASE(nearest_eq, nearest_eq) else PSACREF_ANSV_CASE(nearest_eq, furthest_eq)
    else PSACREF_ANSV_CASE(furthest_eq, nearest_sm) else PSACREF_ANSV_CASE(furthest_eq, nearest_eq) else PSACREF_ANSV_CASE(furthest_eq, furthest_eq)
    else return -1;
#undef PSACREF_ANSV_CASE
    for (size_t i = 0; i < n; ++i) { left_out[i] = l[i]; right_out[i] = r[i]; }
    return 0;
}

/* construct (u64, LCP) + construct_suffix_tree (suffix_tree.hpp:413-499) at np=1.
 * nodes_out must hold (sigma+1)*n entries; returns sigma+1 (row width) or <0. */
long psacref_suffix_tree(const char* text, size_t n, uint64_t* nodes_out, size_t nodes_cap) {
    try {
        cerr_mute mute(!g_verbose);
        mxx::comm c;
        suffix_array<char, size_t, true> sa(c);
        sa.construct(text, text + n);
        std::vector<size_t> nodes = construct_suffix_tree(sa, text, text + n, c);
        if (nodes.size() > nodes_cap) return -2;
        for (size_t i = 0; i < nodes.size(); ++i) nodes_out[i] = nodes[i];
        return (long)(sa.alpha.sigma() + 1);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_suffix_tree: %s\n", e.what());
        return -10;
    }
}

/* File formats of the reference's own suffix_array::write / read (include/suffix_array.hpp:232-265; .sa / .lcp raw index_t
 * through mxx::coll_file, .alpha through alphabet::write): used by tests/test_fileio.py to cross-check psac_b200/fileio.py and the
 * C++ mirror's write / read against files the UNMODIFIED reference wrote, and to let the reference read files written here. */
int psacref_write(const char* text, size_t n, int index_bytes, const char* basename) {
    try {
        cerr_mute mute(!g_verbose);
        mxx::comm c;
        if (index_bytes == 4) {
            suffix_array<char, uint32_t, true> sa(c);
            sa.construct(text, text + n);
            sa.write(basename);
        } else {
            suffix_array<char, uint64_t, true> sa(c);
            sa.construct(text, text + n);
            sa.write(basename);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_write: %s\n", e.what());
        return -10;
    }
}

/* reads <basename>.sa / .lcp / .alpha with the reference's read(); returns n (or < 0), fills sa / lcp (cap elements of
 * index_bytes each) and lut256[c] = alpha.encode(c) */
long psacref_read(const char* basename, int index_bytes, void* sa_out, void* lcp_out, size_t cap, uint8_t* lut256, unsigned* sigma) {
    try {
        cerr_mute mute(!g_verbose);
        mxx::comm c;
        if (index_bytes == 4) {
            suffix_array<char, uint32_t, true> sa(c);
            sa.read(basename);
            if (sa.local_SA.size() > cap) return -2;
            std::memcpy(sa_out, sa.local_SA.data(), sa.local_SA.size() * 4);
            std::memcpy(lcp_out, sa.local_LCP.data(), sa.local_LCP.size() * 4);
            for (int ch = 0; ch < 256; ++ch) lut256[ch] = sa.alpha.encode((char)ch);
            *sigma = sa.alpha.sigma();
            return (long)sa.n;
        } else {
            suffix_array<char, uint64_t, true> sa(c);
            sa.read(basename);
            if (sa.local_SA.size() > cap) return -2;
            std::memcpy(sa_out, sa.local_SA.data(), sa.local_SA.size() * 8);
            std::memcpy(lcp_out, sa.local_LCP.data(), sa.local_LCP.size() * 8);
            for (int ch = 0; ch < 256; ++ch) lut256[ch] = sa.alpha.encode((char)ch);
            *sigma = sa.alpha.sigma();
            return (long)sa.n;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_read: %s\n", e.what());
        return -10;
    }
}

/* suffix_array<char, uint64_t, true, true>::local_Lc: the left-branching characters (suffix_array.hpp:212, 1365-1383, 1485-1495) */
int psacref_lc(const char* text, size_t n, unsigned k, char* lc_out) {
    try {
        cerr_mute mute(!g_verbose);
        mxx::comm c;
        suffix_array<char, uint64_t, true, true> sa(c);
        sa.construct(text, text + n, true, k);
        if (sa.local_Lc.size() != n) return -3;
        std::memcpy(lc_out, sa.local_Lc.data(), n);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_lc: %s\n", e.what());
        return -10;
    }
}

/* rand_dna(size, seed) exactly as the reference's tests generate inputs (alphabet.hpp:37-45; glibc rand) */
int psacref_rand_dna(size_t n, int seed, char* out) {
    std::string s = rand_dna(n, seed);
    std::memcpy(out, s.data(), n);
    return 0;
}

long psacref_construct_ss(const char* flat, size_t len, char sep, const char* alpha_chars, size_t n_alpha, int index_bytes, void* sa_out, void* isa_out,
                          void* lcp_out, size_t cap) {
    try {
        if (index_bytes == 8) return run_construct_ss<uint64_t>(flat, len, sep, alpha_chars, n_alpha, sa_out, isa_out, lcp_out, cap);
        if (index_bytes == 4) return run_construct_ss<uint32_t>(flat, len, sep, alpha_chars, n_alpha, sa_out, isa_out, lcp_out, cap);
        return -1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_construct_ss: %s\n", e.what());
        return -10;
    }
}

int psacref_construct_int(const int* text, size_t n, int index_bytes, int want_lcp, void* sa_out, void* isa_out, void* lcp_out) {
    try {
        if (index_bytes == 4) return want_lcp ? run_construct_int<uint32_t, true>(text, n, sa_out, isa_out, lcp_out) : run_construct_int<uint32_t, false>(text, n, sa_out, isa_out, lcp_out);
        if (index_bytes == 8) return want_lcp ? run_construct_int<uint64_t, true>(text, n, sa_out, isa_out, lcp_out) : run_construct_int<uint64_t, false>(text, n, sa_out, isa_out, lcp_out);
        return -1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "psacref_construct_int: %s\n", e.what());
        return -10;
    }
}

} // extern "C"

namespace { struct env_holder { mxx::env e; }; env_holder* g_env = nullptr; }
extern "C" void psacref_init() { if (!g_env) g_env = new env_holder(); }
