"""Generalized suffix array of a string set (reference construct_ss, include/suffix_array.hpp:269-363).

CPU part: the plain-C restatement (oracle/psac_oracle.c oracle_construct_ss) against the reference's golden vector
(test/test_gsa.cpp:97-98), its closed-form expectations (:31-66), fixtures generated from the unmodified reference
(tests/golden/gsa_*.npz, tests/golden/make_golden.py) and, where oracle/_ref exists, the unmodified reference itself.
GPU part: the CUDA path through the C ABI (psacb200_construct_ss) against the same, bit-exact."""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from psac_b200 import textgen as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def repeat_inc_flat(seq, reps):
    """test/test_gsa.cpp:23-29 repeat_inc_seq + flatten_strings"""
    return b"$".join(seq * (i + 1) for i in range(reps))


def repeat_inc_gsa(slen, reps):
    """closed form of the expected GSA, test/test_gsa.cpp:31-49"""
    m = reps * (reps + 1) // 2
    gsa = np.zeros(slen * m, np.uint64)
    for i in range(slen):
        o = i * m
        for j in range(reps):
            gsa[o] = i + slen * (j * (j + 1)) // 2
            o += 1
            for k in range(j + 2, reps + 1):
                gsa[o] = gsa[o - 1] + np.uint64(k * slen)
                o += 1
    return gsa


def repeat_inc_glcp(slen, reps):
    """closed form of the expected LCP, test/test_gsa.cpp:51-66"""
    m = reps * (reps + 1) // 2
    lcp = np.zeros(slen * m, np.uint64)
    for i in range(slen):
        o = i * m
        lcp[o] = 0
        o += 1
        for j in range(1, reps):
            for _ in range(reps + 1 - j):
                lcp[o] = j * slen - i
                o += 1
    return lcp


INC_CASES = [(b"ab", 3), (b"abc", 3), (b"a", 20), (b"abc", 10), (b"abcdef", 50)]  # test_gsa.cpp:150-168


def _same(a, b):
    return all(np.array_equal(np.asarray(a[k], np.uint64), np.asarray(b[k], np.uint64)) for k in ("sa", "isa", "lcp"))


# ------------------------------------------------------------------------------------------------ oracle (CPU)
def test_port_simple_tiny_golden():
    r = O.construct_ss(b"abab$baba")
    assert r["sa"].tolist() == [7, 2, 5, 0, 3, 6, 1, 4]  # test_gsa.cpp:97
    assert r["lcp"].tolist() == [0, 1, 2, 3, 0, 1, 2, 3]  # test_gsa.cpp:98
    assert (r["isa"][r["sa"].astype(np.int64)] == np.arange(8)).all()


@pytest.mark.parametrize("seq,reps", INC_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_port_inc_repeats_closed_forms(seq, reps, bits):
    r = O.construct_ss(repeat_inc_flat(seq, reps), index_bits=bits)
    assert np.array_equal(r["sa"], repeat_inc_gsa(len(seq), reps))
    assert np.array_equal(r["lcp"], repeat_inc_glcp(len(seq), reps))


def test_port_matches_reference_fixtures():
    files = sorted(glob.glob(os.path.join(GOLD, "gsa_*.npz")))
    assert len(files) >= 6
    for f in files:
        g = np.load(f)
        r = O.construct_ss(g["flat"], index_bits=8 * int(g["index_bytes"]))
        assert _same(r, g), f
        assert _same(O.gsa_naive(g["flat"]), g), f  # the fixtures satisfy the definition


def test_port_edge_cases():
    for flat in (b"a", b"$a$", b"a$a$a", b"$$$ab$$", b"ab", b"b$a", b"aaaa$aaaa$aaaa"):
        r, d = O.construct_ss(flat), O.gsa_naive(flat)
        assert _same(r, d), flat
    assert O.construct_ss(b"$$$")["n"] == 0


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libpsacref.so not built")
def test_port_vs_unmodified_reference_randomised():
    rng = np.random.default_rng(5)
    for t in range(120):
        al = b"acgtn"[: int(rng.integers(1, 6))]
        flat = G.random_stringset(int(rng.integers(1, 25)), int(rng.integers(2, 90)), 1000 + t, alphabet=al, repeat_unit=(3 if t % 3 == 0 else 0))
        if t % 4 == 0:
            flat = np.frombuffer(b"$" + flat.tobytes().replace(b"$", b"$$") + b"$", np.uint8)
        if int((flat != ord("$")).sum()) < 2:
            continue
        d = O.gsa_naive(flat)
        for ib in (4, 8):
            assert _same(O.construct_ss(flat, index_bits=8 * ib), d), (t, ib)
            assert _same(O.ref_construct_ss(flat, index_bytes=ib), d), (t, ib)


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
@pytest.fixture(scope="module")
def eng():
    from psac_b200 import api
    e = api.Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_simple_tiny_and_closed_forms(eng):
    r = eng.construct_ss(b"abab$baba")
    assert r["n"] == 8 and r["sa"].tolist() == [7, 2, 5, 0, 3, 6, 1, 4] and r["lcp"].tolist() == [0, 1, 2, 3, 0, 1, 2, 3]
    for seq, reps in INC_CASES:
        for ib in (4, 8):
            r = eng.construct_ss(repeat_inc_flat(seq, reps), index_bytes=ib)
            assert np.array_equal(r["sa"].astype(np.uint64), repeat_inc_gsa(len(seq), reps)), (seq, reps, ib)
            assert np.array_equal(r["lcp"].astype(np.uint64), repeat_inc_glcp(len(seq), reps)), (seq, reps, ib)
            assert (r["isa"][r["sa"].astype(np.int64)] == np.arange(r["n"])).all()


@pytest.mark.gpu
def test_gpu_matches_reference_fixtures(eng):
    for f in sorted(glob.glob(os.path.join(GOLD, "gsa_*.npz"))):
        g = np.load(f)
        r = eng.construct_ss(g["flat"], index_bytes=int(g["index_bytes"]))
        assert r["n"] == g["sa"].size and _same(r, g), f


@pytest.mark.gpu
def test_gpu_edge_cases(eng):
    for flat in (b"a", b"$a$", b"a$a$a", b"$$$ab$$", b"ab", b"b$a", b"aaaa$aaaa$aaaa", b"z" * 300, b"$".join([b"ab" * 40] * 50)):
        for ib in (4, 8):
            r, d = eng.construct_ss(flat, index_bytes=ib), O.gsa_naive(flat)
            assert r["n"] == d["n"] and _same(r, d), (flat[:20], ib)
    assert eng.construct_ss(b"$$$")["n"] == 0
    assert eng.construct_ss(b"")["n"] == 0
    r = eng.construct_ss(b"ab$ba", want_lcp=False, want_isa=False)  # SA only
    assert r["sa"].tolist() == [3, 0, 1, 2] and r["lcp"] is None and r["isa"] is None


@pytest.mark.gpu
def test_gpu_matches_oracle_randomised(eng):
    rng = np.random.default_rng(9)
    for t in range(40):
        al = [b"a", b"ab", b"ACGT", b"ACGTN", bytes(range(65, 65 + 20)), bytes(range(1, 36)) + bytes(range(37, 256))][t % 6]
        flat = G.random_stringset(int(rng.integers(1, 400)), int(rng.integers(2, 300)), 2000 + t, alphabet=al, repeat_unit=(4 if t % 2 else 0))
        exp = O.construct_ss(flat)
        for ib in (4, 8):
            r = eng.construct_ss(flat, index_bytes=ib)
            assert r["n"] == exp["n"] and _same(r, exp), (t, ib)


@pytest.mark.gpu
def test_gpu_other_separator_and_user_alphabet(eng):
    flat = G.random_stringset(200, 100, 77, alphabet=b"xyz", sep=b"\n")
    exp = O.construct_ss(flat, sep=10)
    r = eng.construct_ss(flat, sep=10)
    assert _same(r, exp)
    # a caller-supplied alphabet: only the ORDER of the codes matters (here: reversed byte order)
    lut = np.zeros(256, np.uint8)
    lut[ord("x")], lut[ord("y")], lut[ord("z")] = 3, 2, 1
    r2 = eng.construct_ss(flat, sep=10, lut=lut)
    swapped = flat.copy()
    swapped[flat == ord("x")] = ord("z")
    swapped[flat == ord("z")] = ord("x")
    assert _same(r2, O.construct_ss(swapped, sep=10))


@pytest.mark.gpu
def test_gpu_many_reads_properties(eng):
    """2^22 characters of short DNA reads (the shape of a sequencing read set): order, ties by position and LCP are checked
    by direct comparison on a sample of neighbours, the permutation on all of them."""
    flat = G.random_stringset(40000, 200, 5, alphabet=b"ACGT")
    r = eng.construct_ss(flat, index_bytes=4)
    n = r["n"]
    assert n == int((flat != ord("$")).sum())
    assert np.array_equal(np.sort(r["sa"]), np.arange(n, dtype=np.uint32))
    assert (r["isa"][r["sa"]] == np.arange(n)).all()
    cat = flat[flat != ord("$")].tobytes()
    ends = np.zeros(n, np.int64)  # exclusive end of every position's string, without separators
    pos = 0
    for s in flat.tobytes().split(b"$"):
        ends[pos:pos + len(s)] = pos + len(s)
        pos += len(s)
    rng = np.random.default_rng(1)
    for q in rng.integers(1, n, 4000):
        a, b = int(r["sa"][q - 1]), int(r["sa"][q])
        sa_, sb_ = cat[a:ends[a]], cat[b:ends[b]]
        assert sa_ < sb_ or (sa_ == sb_ and a < b)
        l = 0
        while l < len(sa_) and l < len(sb_) and sa_[l] == sb_[l]:
            l += 1
        assert int(r["lcp"][q]) == l
    assert r["lcp"][0] == 0


@pytest.mark.gpu
def test_gpu_python_mirror(eng):
    from psac_b200 import api
    sa = api.SuffixArray(8, construct_lcp=True, engine=eng).construct_ss(b"abab$baba")
    assert sa.n == 8 and sa.local_SA.tolist() == [7, 2, 5, 0, 3, 6, 1, 4] and sa.local_LCP.tolist() == [0, 1, 2, 3, 0, 1, 2, 3]
