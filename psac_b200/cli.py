"""``python -m psac_b200.cli`` -- the flags of the reference's ``psac`` tool (reference src/psac.cpp:56-156) on one B200.

    -f FILE | -r SIZE [-s SEED]   input text (file, or random DNA like the reference's rand_dna)
    -l                            also build the LCP array          (suffix_array<char, uint64_t, true>)
    -t                            also build the suffix tree        (construct_suffix_tree)
    -c                            check the result (ISA is the inverse of SA, suffixes in order, LCP by direct comparison
                                  on a sample -- the reference's gl_check_correct / d_check_sa do the full O(n) pass)
    -o BASE                       write BASE.sa64 (and BASE.lcp64 with -l), raw little-endian uint64 like the reference
Timing lines go to stderr in the reference's wording ("PSAC time: ... ms").  There is no CPU path: without a GPU the
tool exits with the library's error.
"""
import argparse
import sys
import time

import numpy as np

from . import api, fileio, textgen


def _check(text, sa, isa, lcp, samples=20000, seed=1):
    n = text.size
    if not (sa[isa.astype(np.int64)] == np.arange(n, dtype=sa.dtype)).all():
        return "ISA is not the inverse of SA"
    rng = np.random.default_rng(seed)
    for p in rng.integers(1, n, size=min(samples, max(n - 1, 0))):
        a, b = int(sa[p - 1]), int(sa[p])
        l = 0
        while a + l < n and b + l < n and text[a + l] == text[b + l]:
            l += 1
        if not (a + l == n or (b + l < n and text[a + l] < text[b + l])):
            return "suffixes %d and %d are out of order" % (a, b)
        if lcp is not None and int(lcp[p]) != l:
            return "LCP[%d] = %d, expected %d" % (p, int(lcp[p]), l)
    return None


def main(argv=None):
    ap = argparse.ArgumentParser(prog="psac_b200.cli", description="Suffix array and LCP construction on a B200 (flags of the reference's psac tool).")
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument("-f", "--file", help="Input filename.")
    g.add_argument("-r", "--random", type=int, help="Random input size")
    ap.add_argument("-o", "--outfile", default="", help="Output file base name.")
    ap.add_argument("-s", "--seed", type=int, default=0, help="Sets the seed for the random input generation")
    ap.add_argument("-l", "--lcp", action="store_true", help="Construct the LCP alongside the SA.")
    ap.add_argument("-t", "--tree", action="store_true", help="Construct the Suffix Tree structure.")
    ap.add_argument("-c", "--check", action="store_true", help="Check correctness of SA (and LCP).")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    text = fileio.file_block_decompose(args.file) if args.file else textgen.random_dna(args.random, args.seed)
    want_lcp = args.lcp or args.tree
    try:
        eng = api.Engine(args.device)
        t0 = time.perf_counter()
        r = eng.construct(text, 8, want_lcp)  # the reference tool is fixed to index_t = uint64_t (psac.cpp:54)
        t1 = time.perf_counter()
        print("PSAC time: %.3f ms" % ((t1 - t0) * 1e3), file=sys.stderr)
        if args.tree:
            nodes = eng.suffix_tree(text, r["sa"], r["lcp"])
            print("ST time: %.3f ms" % ((time.perf_counter() - t1) * 1e3), file=sys.stderr)
            print("suffix tree: %d x %d child table, %d cells used" % (nodes.shape[0], nodes.shape[1], int((nodes != 0).sum())), file=sys.stderr)
    except api.PsacError as e:
        print("psac_b200: %s" % e, file=sys.stderr)
        return 2
    if args.check:
        err = _check(text, r["sa"], r["isa"], r["lcp"] if args.lcp else None)
        print("check: %s" % (err or "ok"), file=sys.stderr)
        if err:
            return 1
    if args.outfile:
        fileio.write_psac_cli_output(args.outfile, r["sa"], r["lcp"] if args.lcp else None)
    return 0


if __name__ == "__main__":
    sys.exit(main())
