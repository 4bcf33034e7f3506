// psacb200::suffix_array over SEVERAL GPUs of one box: the object a p = 1 caller of the reference class binds
// (psacb200::comm(device, n_gpus) -> psacb200_multi_construct).  The result must equal the one-GPU result element for
// element (SA, ISA, LCP), on a text large enough to take the sharded path and on a small one (replicated path).
// usage: test_shim_multi <n_gpus>.  Exit 0 = ok, 3 = no CUDA device (the shim threw), else failure.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "psacb200/suffix_array.hpp"

static int fail(const char* what) {
    std::fprintf(stderr, "FAIL: %s\n", what);
    return 1;
}

int main(int argc, char** argv) {
    const int gpus = argc > 1 ? std::atoi(argv[1]) : 2;
    try {
        for (size_t n : {(size_t)1000, (size_t)(gpus << 18) + 7}) {
            std::vector<char> t(n);
            uint64_t x = 88172645463325252ull + n;
            for (auto& ch : t) {
                x ^= x << 13, x ^= x >> 7, x ^= x << 17;
                ch = "ACGT"[x & 3];
            }
            psacb200::suffix_array<char, uint64_t, true> one(psacb200::comm(0, 1));
            one.construct(t.begin(), t.end());
            psacb200::suffix_array<char, uint64_t, true> many(psacb200::comm(0, gpus));
            many.construct(t.begin(), t.end());
            if (many.p != 1 || many.n != n || many.local_SA.size() != n) return fail("sizes");
            if (many.local_SA != one.local_SA) return fail("SA differs from the one-GPU result");
            if (many.local_B != one.local_B) return fail("ISA differs from the one-GPU result");
            if (many.local_LCP != one.local_LCP) return fail("LCP differs from the one-GPU result");
            many.construct(t.begin(), t.end(), true, 5);  // again on the same object, short first key
            if (many.local_SA != one.local_SA || many.local_LCP != one.local_LCP) return fail("k = 5 differs");
            psacb200::suffix_array<char, uint32_t, false> m32(psacb200::comm(0, gpus));
            m32.construct(t.begin(), t.end());
            for (size_t i = 0; i < n; ++i)
                if (m32.local_SA[i] != one.local_SA[i]) return fail("32-bit index differs");
        }
    } catch (const std::runtime_error& e) {
        std::fprintf(stderr, "runtime_error: %s\n", e.what());
        return std::string(e.what()).find("no CUDA device") != std::string::npos ? 3 : 2;
    }
    std::printf("shim multi ok (%d GPUs)\n", gpus);
    return 0;
}
