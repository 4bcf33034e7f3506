"""Texts over wide characters (reference suffix_array<int, index_t, LCP> with int_alphabet, include/alphabet.hpp:355-513;
test/test_psac.cpp:277-304 "IntAlphabetMiss").  CPU: the restatement (rank the values, then the byte construction) against the
golden vector and the unmodified reference.  GPU: psacb200_construct_wide through the C ABI against the restatement."""
import numpy as np
import pytest

from oracle import pyoracle as O

MISS = np.array([128, 3, 12345678, 12345678, 3, 12345678, 12345678, 3, 66000, 66000, 3], np.int32)  # test_psac.cpp:287
MISS_SA = [10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2]                                                           # test_psac.cpp:286


def _texts():
    rng = np.random.default_rng(7)
    for lo, hi, dt in ((0, 2 ** 15, np.int32), (-1000, 1000, np.int32), (-2 ** 31, 2 ** 31 - 1, np.int32), (0, 2 ** 32 - 1, np.uint32),
                       (-2 ** 15, 2 ** 15 - 1, np.int16), (0, 2 ** 16 - 1, np.uint16)):
        for nd in (1, 2, 7, 255):
            al = np.unique(rng.integers(lo, hi, size=nd, endpoint=True))
            for n in (1, 2, 50, 4097, 100003):
                yield al[rng.integers(0, al.size, n)].astype(dt)


def test_port_golden_and_reference():
    assert O.construct_wide(MISS, 32, True)["sa"].tolist() == MISS_SA
    if not O.have_ref():
        pytest.skip("oracle/_ref/libpsacref.so not built")
    assert O.ref_construct_int(MISS, 4, True)["sa"].tolist() == MISS_SA
    rng = np.random.default_rng(3)
    for t in range(40):
        # 64-bit index: any values; 32-bit index: <= 16 bits per character (beyond that the reference drops to k = 1 and
        # copies raw sign-extended characters, include/kmer.hpp:196-199, which misorders them -- not reproduced)
        for ib, lo, hi in ((8, -2 ** 31, 2 ** 31 - 1), (8, 0, 2 ** 24), (4, 0, 2 ** 15), (4, -1000, 1000)):
            al = rng.integers(lo, hi, size=int(rng.integers(1, 40)))
            text = al[rng.integers(0, al.size, int(rng.integers(2, 2000)))].astype(np.int32)
            r, p = O.ref_construct_int(text, ib, True), O.construct_wide(text, 64, True)
            assert all(np.array_equal(r[k].astype(np.uint64), p[k]) for k in ("sa", "isa", "lcp")), (t, ib, lo, hi)


@pytest.fixture(scope="module")
def eng():
    from psac_b200 import api
    e = api.Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_int_alphabet_miss_golden(eng):
    r = eng.construct_wide(MISS, 4, True)
    assert r["sa"].tolist() == MISS_SA and r["lcp"].tolist() == [0, 1, 1, 4, 0, 0, 1, 0, 2, 1, 3]
    assert r["distinct"].tolist() == [3, 128, 66000, 12345678]


@pytest.mark.gpu
def test_gpu_wide_matches_port(eng):
    for text in _texts():
        for ib in (4, 8):
            r, p = eng.construct_wide(text, ib, True), O.construct_wide(text, 64, True)
            assert all(np.array_equal(r[k].astype(np.uint64), p[k]) for k in ("sa", "isa", "lcp")), (text.dtype, text.size, ib)
            assert np.array_equal(r["distinct"], np.unique(text).astype(np.int64))


@pytest.mark.gpu
def test_gpu_wide_rejects_more_than_255_values(eng):
    from psac_b200 import api
    with pytest.raises(api.PsacError) as ei:
        eng.construct_wide(np.arange(5000, dtype=np.int32), 8)
    assert "255" in str(ei.value)
    with pytest.raises(api.PsacError):
        eng.construct_wide(np.arange(256, dtype=np.uint16), 4)
    assert eng.construct_wide(np.arange(255, dtype=np.uint16), 4)["sa"].tolist() == list(range(255))
