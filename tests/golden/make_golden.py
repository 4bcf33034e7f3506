"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, np=1).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The vectors are committed; the GPU box and CI only read them.  Inputs come from
psac_b200.textgen (own counter-based PRNG) except `refdna_*`, which use the reference's
glibc-rand generator rand_dna(size, seed) with the exact sizes/seeds of test/test_psac.cpp.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402
from psac_b200 import textgen as G  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(name, text, index_bytes, want_lcp, k=0, fast=True):
    text = np.ascontiguousarray(np.frombuffer(text, np.uint8) if isinstance(text, bytes) else text, np.uint8)
    r = O.ref_construct(text, index_bytes, want_lcp, k, fast)
    d = dict(text=text, sa=r["sa"], isa=r["isa"], index_bytes=np.int32(index_bytes), k=np.int32(k))
    if want_lcp:
        d["lcp"] = r["lcp"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, text.size, "ok")


def main():
    case("mississippi_u32", b"mississippi", 4, False)
    case("mississippi_u64_lcp", b"mississippi", 8, True)
    case("dna4096_u32", G.random_dna(4096, 11), 4, False)
    case("dna20000_u64_lcp", G.random_dna(20000, 12), 8, True)
    case("dna9000_u64_lcp_k3", G.random_dna(9000, 13), 8, True, k=3)          # bucket chasing + RMQ LCP path
    case("dna7000_u32_nofast_k2", G.random_dna(7000, 14), 4, False, k=2, fast=False)
    case("refdna9_u32", O.ref_rand_dna(9, 13), 4, False)                        # test_psac.cpp:226-248
    case("refdna2000_u64_lcp", O.ref_rand_dna(2000, 23), 8, True)              # shape of test_psac.cpp:250-274
    case("repeats600_u64_lcp", G.repeats_text(600, 1), 8, True)                # shape of test_psac.cpp:178-224
    case("abc97_u64_lcp", G.periodic_text(b"abc", 97), 8, True)                # test_suffixtree.cpp:131-162
    case("aaaa_u32_lcp", G.periodic_text(b"a", 777), 4, True)
    case("bytes255_u64", np.minimum(G.random_bytes(5000, 4), 254).astype(np.uint8), 8, False)
    case("bytes256_u64", G.random_bytes_config4(6000, 4), 8, False)            # sigma=256 LUT wrap
    t = G.random_bytes(300, 5)
    t[-1] = 7
    case("bytes_small_u32_lcp", t, 4, True)
    # k-mer generation vectors (kmer.hpp:204-224) straight from the reference
    txt = G.random_dna(300, 21)
    np.savez_compressed(os.path.join(OUT, "kmers_dna300.npz"), text=txt,
                        k10_u32=O.ref_kmer_generation(txt, 4, 10), k21_u64=O.ref_kmer_generation(txt, 8, 21),
                        k3_u32=O.ref_kmer_generation(txt, 4, 3), k1_u64=O.ref_kmer_generation(txt, 8, 1))
    # suffix tree of mississippi from the reference (equals test_suffixtree.cpp:66-79)
    np.savez_compressed(os.path.join(OUT, "stree_mississippi.npz"), nodes=O.ref_suffix_tree(b"mississippi"))
    # ANSV, all 3x3 mode combinations on a small array with ties (shape of test_ansv.cpp:316-317)
    vals = (G.splitmix64(7, np.arange(137, dtype=np.uint64)) % np.uint64(10)).astype(np.uint64)
    d = dict(vals=vals)
    for lt in range(3):
        for rt in range(3):
            l, r = O.ref_ansv(vals, lt, rt, 2 ** 64 - 1)
            d["l_%d_%d" % (lt, rt)] = l
            d["r_%d_%d" % (lt, rt)] = r
    np.savez_compressed(os.path.join(OUT, "ansv137.npz"), **d)
    gsa_cases()


def gsa_case(name, flat, index_bytes):
    """generalized SA / LCP of a string set from the reference's construct_ss (suffix_array.hpp:269-363)"""
    flat = np.ascontiguousarray(np.frombuffer(flat, np.uint8) if isinstance(flat, bytes) else flat, np.uint8)
    r = O.ref_construct_ss(flat, ord("$"), index_bytes)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), flat=flat, sa=r["sa"], isa=r["isa"], lcp=r["lcp"], index_bytes=np.int32(index_bytes))
    print(name, flat.size, r["n"], "ok")


def gsa_cases():
    gsa_case("gsa_tiny_u64", b"abab$baba", 8)                                                          # test_gsa.cpp:71-103
    gsa_case("gsa_incabc10_u64", b"$".join(b"abc" * (i + 1) for i in range(10)), 8)                     # test_gsa.cpp:160-163
    gsa_case("gsa_incaf50_u32", b"$".join(b"abcdef" * (i + 1) for i in range(50)), 4)                   # test_gsa.cpp:165-168
    gsa_case("gsa_dna_reads_u64", G.random_stringset(300, 150, 31), 8)
    gsa_case("gsa_dna_repeats_u32", G.random_stringset(200, 400, 32, repeat_unit=5), 4)
    gsa_case("gsa_ragged_seps_u64", b"$$" + G.random_stringset(40, 30, 33, alphabet=b"ab").tobytes().replace(b"$", b"$$$") + b"$", 8)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gsa":
        gsa_cases()
    else:
        main()
