// Exercises include/psacb200/suffix_array.hpp the way the reference's own tests use its class
// (reference test/test_psac.cpp:101-129 "Mississippi", :131-176 "RandAll" construct / construct_arr).
// Exit codes: 0 = all checks passed on a GPU; 3 = no CUDA device and the shim threw std::runtime_error as it must
// (there is no CPU fallback); anything else = failure.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "psacb200/suffix_array.hpp"

static int fail(const char* what) {
    std::fprintf(stderr, "FAIL: %s\n", what);
    return 1;
}

int main() {
    const std::string s = "mississippi";
    const size_t golden[11] = {10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2};  // test_psac.cpp:105
    try {
        psacb200::comm c(0);
        psacb200::suffix_array<char, uint64_t, true> sa(c);
        sa.construct(s.begin(), s.end());
        if (sa.n != 11 || sa.local_size != 11 || sa.p != 1) return fail("sizes");
        for (size_t i = 0; i < 11; ++i)
            if (sa.local_SA[i] != golden[i]) return fail("mississippi SA");
        for (size_t i = 0; i < 11; ++i)
            if (sa.local_B[sa.local_SA[i]] != i) return fail("ISA is not the inverse of SA");
        const uint64_t lcp[11] = {0, 1, 1, 4, 0, 0, 1, 0, 2, 1, 3};
        for (size_t i = 0; i < 11; ++i)
            if (sa.local_LCP[i] != lcp[i]) return fail("mississippi LCP");
        if (sa.alpha.sigma() != 4 || sa.alpha.bits_per_char() != 3 || sa.alpha.encode('i') != 1 || sa.alpha.encode('s') != 4) return fail("alphabet");

        // repeated construct on one object with different k, like test_psac.cpp:148-171
        std::vector<char> t(20000);
        uint64_t x = 88172645463325252ull;
        for (auto& ch : t) {
            x ^= x << 13, x ^= x >> 7, x ^= x << 17;
            ch = "ACGT"[x & 3];
        }
        psacb200::suffix_array<char, uint32_t, false> sb(c);
        sb.construct(t.begin(), t.end());
        std::vector<uint32_t> ref_sa = sb.local_SA;
        sb.construct(t.begin(), t.end(), true, 3);
        if (sb.local_SA != ref_sa) return fail("k=3 differs");
        sb.construct(t.begin(), t.end(), false, 2);
        if (sb.local_SA != ref_sa) return fail("fast_resolval=false, k=2 differs");
        sb.construct_arr<3>(t.begin(), t.end());
        if (sb.local_SA != ref_sa || !sb.local_LCP.empty()) return fail("construct_arr<3> differs");
        for (size_t i = 1; i < t.size(); ++i) {  // order check on the text itself
            const uint32_t a = ref_sa[i - 1], b = ref_sa[i];
            const size_t la = t.size() - a, lb = t.size() - b, m = la < lb ? la : lb;
            const int cmp = std::memcmp(t.data() + a, t.data() + b, m);
            if (cmp > 0 || (cmp == 0 && la > lb)) return fail("SA not sorted");
        }
        // left-branching characters (_CONSTRUCT_LC, suffix_array.hpp:212): Lc[i] = S[SA[i-1] + LCP[i]], '\0' past the end
        psacb200::suffix_array<char, uint64_t, true, true> sl(c);
        sl.construct(s.begin(), s.end());
        if (sl.local_Lc.size() != 11) return fail("Lc size");
        for (size_t i = 0; i < 11; ++i) {
            const size_t g = i ? sl.local_SA[i - 1] + sl.local_LCP[i] : 11;
            const char want = g < 11 ? s[g] : '\0';
            if (sl.local_Lc[i] != want) return fail("Lc");
        }
        // user-supplied alphabet overload (suffix_array.hpp:365-366)
        psacb200::suffix_array<char, uint64_t, true> sc(c);
        sc.construct(s.begin(), s.end(), true, sa.alpha, 2);
        for (size_t i = 0; i < 11; ++i)
            if (sc.local_SA[i] != golden[i] || sc.local_LCP[i] != lcp[i]) return fail("alphabet overload");
        // generalized suffix array of a string set (reference test/test_gsa.cpp:71-103 "SimpleTiny")
        {
            const std::string flat = "abab$baba";
            psacb200::simple_dstringset ss(flat.begin(), flat.end(), c);
            psacb200::alphabet a = psacb200::alphabet::from_string("ab", c);
            psacb200::suffix_array<char, uint64_t, true> sg(c);
            sg.construct_ss(ss, a);
            const uint64_t ex_gsa[8] = {7, 2, 5, 0, 3, 6, 1, 4}, ex_lcp[8] = {0, 1, 2, 3, 0, 1, 2, 3};  // test_gsa.cpp:97-98
            if (sg.n != 8 || sg.local_SA.size() != 8 || sg.local_LCP.size() != 8) return fail("GSA sizes");
            for (size_t i = 0; i < 8; ++i)
                if (sg.local_SA[i] != ex_gsa[i] || sg.local_LCP[i] != ex_lcp[i] || sg.local_B[sg.local_SA[i]] != i) return fail("GSA SimpleTiny");
        }
        // wide characters ordered by value (reference test/test_psac.cpp:277-304 "IntAlphabetMiss")
        {
            const std::vector<int> str = {128, 3, 12345678, 12345678, 3, 12345678, 12345678, 3, 66000, 66000, 3};
            psacb200::suffix_array<int, unsigned int, true> si(c);
            si.construct(str.begin(), str.end());
            for (size_t i = 0; i < 11; ++i)
                if (si.local_SA[i] != golden[i] || si.local_LCP[i] != lcp[i]) return fail("IntAlphabetMiss");
            if (si.alpha.min_char != 3 || si.alpha.max_char != 12345678 || si.alpha.bits_per_char() != 24) return fail("int_alphabet");
        }
    } catch (const std::runtime_error& e) {
        if (std::strstr(e.what(), "no CUDA device") || std::strstr(e.what(), "CUDA")) {
            std::fprintf(stderr, "no GPU: %s\n", e.what());
            return 3;
        }
        std::fprintf(stderr, "unexpected error: %s\n", e.what());
        return 2;
    }
    std::puts("shim ok");
    return 0;
}
