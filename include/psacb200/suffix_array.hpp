// psacb200::suffix_array -- C++ host-side mirror of patflick/psac's suffix_array<char_t, index_t, LCP> for the
// construct() / construct_arr<L>() path, forwarding to the C ABI in include/psacb200.h.
//
// Reference class surface mirrored here (reference include/suffix_array.hpp):
//   class template + public members n, local_size, p, local_SA, local_B (= ISA), local_LCP   :169-212
//   init_size(lsize)  -- throws std::runtime_error on a bad block decomposition              :217-228
//   construct(begin, end, fast_resolval = true, k = 0)                                       :469-486
//   construct(begin, end, fast_resolval, alphabet, k)                                        :365-366
//   construct_arr<L>(begin, end, fast_resolval = true)   (SA / ISA only, no LCP :555-567)    :490-641
//   write(basename) / read(basename)   (.sa / .lcp raw index_t arrays, .alpha)               :232-265
// Differences, all outside the hot path: the communicator argument is psacb200::comm, a stand-in for an mxx::comm of
// size 1 (the GPUs of one box are sharded INSIDE the engine, include/psacb200.h psacb200_construct_sharded, so the
// host-facing object always holds the whole arrays); the alphabet is the 256-entry code table
// of the reference's alphabet<char>::mapping_table (include/alphabet.hpp:136,157-164).
// Errors surface as std::runtime_error carrying psacb200_last_error(), like the reference's own throw at :226-227.
// There is no CPU fallback: without libpsacb200.so and a B200 every construct call throws.
#ifndef PSACB200_SUFFIX_ARRAY_HPP
#define PSACB200_SUFFIX_ARRAY_HPP

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../psacb200.h"

namespace psacb200 {

// stand-in for mxx::comm at p = 1 (reference ext/mxx/include/mxx/comm_fwd.hpp:44-).  device = first CUDA ordinal to use;
// n_gpus > 1: the construction is sharded over the devices device .. device + n_gpus - 1 INSIDE the engine
// (psacb200_multi_construct): the object still holds the whole arrays, size() stays 1.
struct comm {
    int device;
    int n_gpus;
    explicit comm(int device_ = 0, int n_gpus_ = 1) : device(device_), n_gpus(n_gpus_ < 1 ? 1 : n_gpus_) {}
    int size() const { return 1; }
    int rank() const { return 0; }
    int gpus() const { return n_gpus; }
    comm copy() const { return *this; }
};

// the reference's alphabet<char>: character -> code table (codes start at 1, 0 = end-of-text padding)
struct alphabet {
    uint8_t mapping_table[256];
    unsigned sigma_;
    unsigned bits_per_char_;
    alphabet() : sigma_(0), bits_per_char_(0) {
        for (int i = 0; i < 256; ++i) mapping_table[i] = 0;
    }
    unsigned sigma() const { return sigma_; }                  // alphabet.hpp:241
    unsigned bits_per_char() const { return bits_per_char_; }  // alphabet.hpp:249
    uint8_t encode(unsigned char c) const { return mapping_table[c]; }  // alphabet.hpp:266-269
    // alphabet.hpp:220-225: the alphabet of the characters of a string (codes 1.. in byte order)
    static alphabet from_string(const std::string& str, const comm& = comm()) {
        alphabet a;
        bool used[256] = {false};
        for (unsigned char c : str) used[c] = true;
        unsigned code = 1;
        for (int i = 0; i < 256; ++i)
            if (used[i]) {
                a.mapping_table[i] = (uint8_t)code++;
                ++a.sigma_;
            }
        while ((1u << a.bits_per_char_) < a.sigma_ + 1) ++a.bits_per_char_;
        return a;
    }
};

// the reference's int_alphabet<IntType> (include/alphabet.hpp:355-502): the characters of a wide-character text are ordered
// by value; the alphabet is the value range [min_char, max_char] that occurs
template <typename IntType>
struct int_alphabet {
    IntType min_char, max_char;
    unsigned long long sigma_;
    unsigned bits_per_char_;
    int_alphabet() : min_char(0), max_char(0), sigma_(0), bits_per_char_(0) {}
    int_alphabet(IntType mn, IntType mx) : min_char(mn), max_char(mx), sigma_((unsigned long long)((long long)mx - (long long)mn) + 1), bits_per_char_(0) {
        while ((1ull << bits_per_char_) < sigma_ + 1) ++bits_per_char_;  // ceillog2(sigma + 1), alphabet.hpp:398
    }
    unsigned long long sigma() const { return sigma_; }
    unsigned bits_per_char() const { return bits_per_char_; }
};

// reference alphabet_helper (alphabet.hpp:509-513); here every character type wider than a byte takes the value-ordered path
template <typename CharType>
struct alphabet_helper {
    using alphabet_type = typename std::conditional<sizeof(CharType) == 1, alphabet, int_alphabet<CharType>>::type;
};

// stand-in for simple_dstringset at p = 1 (reference include/stringset.hpp:33-152): the strings are the maximal runs of
// non-separator characters of a flat text, which is copied once (the reference borrows it)
class simple_dstringset {
public:
    std::vector<uint8_t> flat;
    char sep;
    std::size_t sum_sizes;
    template <typename Iterator>
    simple_dstringset(Iterator begin, Iterator end, const comm& = comm(), char sep_ = '$') : flat(begin, end), sep(sep_), sum_sizes(0) {
        for (uint8_t c : flat) sum_sizes += (c != (uint8_t)sep) ? 1 : 0;
    }
};

template <typename char_t, typename index_t = std::size_t, bool _CONSTRUCT_LCP = false, bool _CONSTRUCT_LC = false>
class suffix_array {
    static_assert(!_CONSTRUCT_LC || _CONSTRUCT_LCP, "the left-branching characters need the LCP array (reference suffix_array.hpp:1365-1383)");
    static_assert(sizeof(char_t) == 1 || sizeof(char_t) == 2 || sizeof(char_t) == 4, "characters of 1, 2 or 4 bytes");
    static_assert(sizeof(char_t) == 1 || !_CONSTRUCT_LC, "left-branching characters are built for 1-byte texts");
    using byte_text = std::integral_constant<bool, sizeof(char_t) == 1>;
    static_assert(sizeof(index_t) == 4 || sizeof(index_t) == 8, "index_t must be a 32- or 64-bit unsigned integer");
    static_assert(std::is_unsigned<index_t>::value, "index_t must be unsigned");

public:
    explicit suffix_array(const psacb200::comm& c) : n(0), local_size(0), comm(c.copy()), p(1), engine_(nullptr), multi_(nullptr) {}
    virtual ~suffix_array() {
        if (multi_)
            psacb200_multi_destroy(multi_);  // (owns its engines)
        else if (engine_)
            psacb200_destroy(engine_);
    }
    suffix_array(const suffix_array&) = delete;
    suffix_array& operator=(const suffix_array&) = delete;

    std::size_t n;
    std::size_t local_size;
    psacb200::comm comm;
    int p;
    using char_type = char_t;
    using alphabet_type = typename psacb200::alphabet_helper<char_t>::alphabet_type;
    alphabet_type alpha;
    std::vector<index_t> local_SA;
    std::vector<index_t> local_B;  // inverse suffix array, 0-based (reference: local_B after :460-464)
    std::vector<index_t> local_LCP;
    std::vector<char_t> local_Lc;  // left-branching characters (reference :212; filled iff _CONSTRUCT_LC)

    // reference :217-228 -- at p = 1 every size is a valid block decomposition
    void init_size(std::size_t lsize) {
        local_size = lsize;
        n = lsize;
        p = comm.size();
    }

    // reference :469-486
    template <typename Iterator>
    void construct(Iterator begin, Iterator end, bool fast_resolval = true, unsigned int k = 0) {
        init_size((std::size_t)std::distance(begin, end));
        construct_impl(begin, end, fast_resolval, k, _CONSTRUCT_LCP, byte_text());
    }

    // reference :365-366 (caller supplies the alphabet; only the ORDER of its codes matters; a wide-character text is ordered
    // by value whatever range the alphabet names)
    template <typename Iterator>
    void construct(Iterator begin, Iterator end, bool fast_resolval, const alphabet_type& alphabet, unsigned int k) {
        init_size((std::size_t)std::distance(begin, end));
        alpha = alphabet;
        construct_alpha_impl(begin, end, fast_resolval, k, byte_text());
    }

    // reference :490-641 -- L-tuple doubling; the final SA / ISA are the same arrays, no LCP is built (:555-567)
    template <std::size_t L, typename Iterator>
    void construct_arr(Iterator begin, Iterator end, bool fast_resolval = true) {
        static_assert(L >= 2, "construct_arr needs L >= 2");
        init_size((std::size_t)std::distance(begin, end));
        construct_impl(begin, end, fast_resolval, 0, false, byte_text());
    }

    // reference :269-363 -- generalized suffix array (+ LCP) of a string set: positions index the concatenation of the strings
    // without separators; identical suffixes of different strings are ordered by position.  Runs on the first GPU.
    void construct_ss(simple_dstringset& ss, const alphabet_type& alphabet) {
        static_assert(sizeof(char_t) == 1, "string sets are 1-byte texts");
        alpha = alphabet;
        ensure_engine();
        const std::size_t cap = ss.flat.size();
        local_SA.resize(cap);
        local_B.resize(cap);
        local_LCP.clear();
        if (_CONSTRUCT_LCP) local_LCP.resize(cap);
        local_Lc.clear();
        uint64_t m = 0;
        check(psacb200_construct_ss(engine_, ss.flat.data(), cap, (uint8_t)ss.sep, (int)sizeof(index_t), _CONSTRUCT_LCP ? PSACB200_LCP : 0u,
                                    alpha.mapping_table, local_SA.data(), local_B.data(), _CONSTRUCT_LCP ? (void*)local_LCP.data() : nullptr, &m));
        local_SA.resize(m);
        local_B.resize(m);
        if (_CONSTRUCT_LCP) local_LCP.resize(m);
        init_size(m);
    }

    // reference :232-243 -- <basename>.sa, .lcp (if built): raw little-endian index_t; .alpha: the used characters
    void write(const std::string& basename) const {
        static_assert(sizeof(char_t) == 1, "the .alpha file format lists 1-byte characters");
        write_array(basename + ".sa", local_SA);
        if (_CONSTRUCT_LCP) write_array(basename + ".lcp", local_LCP);
        std::ofstream f(basename + ".alpha", std::ios::binary);
        for (int c = 0; c < 256; ++c)
            if (alpha.mapping_table[c] != 0 || (alpha.sigma_ == 256 && c == 255)) f.put((char)c);
    }
    // reference :245-265
    void read(const std::string& basename) {
        static_assert(sizeof(char_t) == 1, "the .alpha file format lists 1-byte characters");
        read_array(basename + ".sa", local_SA);
        if (_CONSTRUCT_LCP) {
            read_array(basename + ".lcp", local_LCP);
            if (local_SA.size() != local_LCP.size()) throw std::runtime_error("SA and LCP have to have same size");
        }
        std::ifstream f(basename + ".alpha", std::ios::binary);
        alpha = alphabet_type();
        bool used[256] = {false};
        char c;
        while (f.get(c)) used[(unsigned char)c] = true;
        unsigned code = 1;
        for (int i = 0; i < 256; ++i)
            if (used[i]) {
                alpha.mapping_table[i] = (uint8_t)code++;  // 8-bit table like the reference (alphabet.hpp:136,160)
                ++alpha.sigma_;
            }
        while ((1u << alpha.bits_per_char_) < alpha.sigma_ + 1) ++alpha.bits_per_char_;
        local_B.clear();
        init_size(local_SA.size());
    }

    psacb200_stats stats() const {
        psacb200_stats s{};
        if (engine_) psacb200_get_stats(engine_, &s);
        return s;
    }

private:
    psacb200_engine* engine_;  // the engine of the (first) GPU
    psacb200_multi* multi_;    // several GPUs: the handle that owns all engines
    std::vector<uint8_t> staging_;

    static void write_array(const std::string& filename, const std::vector<index_t>& v) {
        std::ofstream f(filename, std::ios::binary | std::ios::trunc);
        f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(index_t)));
        if (!f) throw std::runtime_error("cannot write " + filename);
    }
    static void read_array(const std::string& filename, std::vector<index_t>& v) {
        std::ifstream f(filename, std::ios::binary | std::ios::ate);
        if (!f) throw std::runtime_error("cannot read " + filename);
        const std::size_t bytes = (std::size_t)f.tellg();
        v.resize(bytes / sizeof(index_t));
        f.seekg(0, std::ios::beg);
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(index_t)));
    }
    static void check(int rc) {
        if (rc != PSACB200_OK) throw std::runtime_error(std::string("psacb200: ") + psacb200_last_error());
    }
    void ensure_engine() {
        if (engine_) return;
        if (comm.gpus() > 1) {
            std::vector<int> devs;
            for (int i = 0; i < comm.gpus(); ++i) devs.push_back(comm.device + i);
            check(psacb200_multi_create(comm.gpus(), devs.data(), &multi_));
            engine_ = psacb200_multi_engine(multi_, 0);
        } else {
            check(psacb200_create(comm.device, &engine_));
        }
    }
    // the C ABI takes one contiguous byte buffer; pointers and vector/string iterators are passed through, anything
    // else is copied once
    template <typename Iterator>
    const uint8_t* contiguous(Iterator begin, Iterator end) {
        if (begin == end) return nullptr;
        if (std::is_pointer<Iterator>::value ||
            std::is_same<Iterator, typename std::vector<char_t>::iterator>::value ||
            std::is_same<Iterator, typename std::vector<char_t>::const_iterator>::value ||
            std::is_same<Iterator, std::string::iterator>::value || std::is_same<Iterator, std::string::const_iterator>::value)
            return reinterpret_cast<const uint8_t*>(&*begin);
        staging_.assign(begin, end);
        return staging_.data();
    }
    // ---- 1-byte characters
    template <typename Iterator>
    void construct_impl(Iterator begin, Iterator end, bool fast_resolval, unsigned k, bool want_lcp, std::true_type) {
        const uint8_t* text = contiguous(begin, end);
        ensure_engine();
        check(psacb200_alphabet(engine_, text, n, alpha.mapping_table, &alpha.sigma_, &alpha.bits_per_char_));
        run(text, fast_resolval, k, nullptr, want_lcp);
    }
    template <typename Iterator>
    void construct_alpha_impl(Iterator begin, Iterator end, bool fast_resolval, unsigned k, std::true_type) {
        run(contiguous(begin, end), fast_resolval, k, alpha.mapping_table, _CONSTRUCT_LCP);
    }
    // ---- 2- and 4-byte characters, ordered by value (reference int_alphabet); runs on the first GPU
    template <typename Iterator>
    void construct_impl(Iterator begin, Iterator end, bool fast_resolval, unsigned k, bool want_lcp, std::false_type) {
        run_wide(begin, end, fast_resolval, k, want_lcp, true);
    }
    template <typename Iterator>
    void construct_alpha_impl(Iterator begin, Iterator end, bool fast_resolval, unsigned k, std::false_type) {
        run_wide(begin, end, fast_resolval, k, _CONSTRUCT_LCP, false);
    }
    template <typename Iterator>
    void run_wide(Iterator begin, Iterator end, bool fast_resolval, unsigned k, bool want_lcp, bool set_alpha) {
        ensure_engine();
        wide_staging_.assign(begin, end);
        local_SA.resize(n);
        local_B.resize(n);
        local_LCP.clear();
        if (want_lcp) local_LCP.resize(n);
        local_Lc.clear();
        int64_t distinct[255];
        uint32_t nd = 0;
        const unsigned flags = (want_lcp ? PSACB200_LCP : 0u) | (fast_resolval ? PSACB200_FAST_RESOLVAL : 0u);
        check(psacb200_construct_wide(engine_, wide_staging_.data(), n, (int)sizeof(char_t), std::is_signed<char_t>::value ? 1 : 0, (int)sizeof(index_t), flags, k,
                                      local_SA.data(), local_B.data(), want_lcp ? (void*)local_LCP.data() : nullptr, distinct, &nd));
        if (set_alpha && nd > 0) alpha = alphabet_type((char_t)distinct[0], (char_t)distinct[nd - 1]);  // from_sequence: [min, max] (alphabet.hpp:404-436)
    }
    std::vector<char_t> wide_staging_;

    void run(const uint8_t* text, bool fast_resolval, unsigned k, const uint8_t* lut, bool want_lcp) {
        ensure_engine();
        local_SA.resize(n);
        local_B.resize(n);
        local_LCP.clear();
        if (want_lcp) local_LCP.resize(n);
        const unsigned flags = (want_lcp ? PSACB200_LCP : 0u) | (fast_resolval ? PSACB200_FAST_RESOLVAL : 0u);
        void* lcp = want_lcp ? (void*)local_LCP.data() : nullptr;
        if (lut)  // (a caller-supplied code order runs on the first GPU)
            check(psacb200_construct_alphabet(engine_, text, n, (int)sizeof(index_t), flags, k, lut, local_SA.data(), local_B.data(), lcp));
        else if (multi_)
            check(psacb200_multi_construct(multi_, text, n, (int)sizeof(index_t), flags, k, local_SA.data(), local_B.data(), lcp));
        else
            check(psacb200_construct(engine_, text, n, (int)sizeof(index_t), flags, k, local_SA.data(), local_B.data(), lcp));
        local_Lc.clear();
        if (_CONSTRUCT_LC && want_lcp && n > 0) {
            local_Lc.resize(n);
            check(psacb200_lc(engine_, text, n, (int)sizeof(index_t), local_SA.data(), local_LCP.data(), reinterpret_cast<uint8_t*>(local_Lc.data())));
        }
    }
};

}  // namespace psacb200
#endif
