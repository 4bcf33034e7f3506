// File formats of the C++ mirror (reference suffix_array::write / read, include/suffix_array.hpp:232-265; its FileIO
// test is test/test_psac.cpp:306-347).  Needs no GPU: the members are filled by hand.
#include <cstdio>
#include <fstream>
#include <string>

#include "psacb200/suffix_array.hpp"

int main(int argc, char** argv) {
    const std::string base = argc > 1 ? argv[1] : "/tmp/psacb200_io";
    psacb200::comm c(0);
    psacb200::suffix_array<char, uint32_t, true> a(c);
    const uint32_t sa[11] = {10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2}, lcp[11] = {0, 1, 1, 4, 0, 0, 1, 0, 2, 1, 3};
    a.local_SA.assign(sa, sa + 11);
    a.local_LCP.assign(lcp, lcp + 11);
    a.init_size(11);
    a.alpha.mapping_table[(unsigned char)'i'] = 1;
    a.alpha.mapping_table[(unsigned char)'m'] = 2;
    a.alpha.mapping_table[(unsigned char)'p'] = 3;
    a.alpha.mapping_table[(unsigned char)'s'] = 4;
    a.alpha.sigma_ = 4;
    a.alpha.bits_per_char_ = 3;
    a.write(base);
    std::ifstream f(base + ".sa", std::ios::binary | std::ios::ate);
    if ((long)f.tellg() != 44) return 1;  // 11 raw uint32
    std::ifstream g(base + ".alpha", std::ios::binary);
    std::string al((std::istreambuf_iterator<char>(g)), std::istreambuf_iterator<char>());
    if (al != "imps") return 2;
    psacb200::suffix_array<char, uint32_t, true> b(c);
    b.read(base);
    if (b.n != 11 || b.local_SA != a.local_SA || b.local_LCP != a.local_LCP) return 3;
    if (b.alpha.sigma() != 4 || b.alpha.bits_per_char() != 3 || b.alpha.encode('s') != 4 || b.alpha.encode('x') != 0) return 4;
    std::puts("io ok");
    return 0;
}
