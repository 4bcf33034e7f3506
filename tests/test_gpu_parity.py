"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle (oracle/), the golden
fixtures generated from the unmodified reference (tests/golden/) and, where the prebuilt oracle/_ref travels, the
unmodified reference itself.  Everything is integer work: the bar is bit-exact."""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from psac_b200 import api
from psac_b200 import textgen as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


# ------------------------------------------------------------------------------------------- radix sort (a6)
@pytest.mark.parametrize("kdt", [np.uint32, np.uint64])
@pytest.mark.parametrize("vdt", [None, np.uint32, np.uint64])
@pytest.mark.parametrize("n", [1, 2, 31, 33, 1000, 4608, 6144, 6145, 100003, (1 << 20) + 7])
def test_sort_pairs_matches_stable_numpy(eng, kdt, vdt, n):
    rng = np.random.default_rng(n * 7 + np.dtype(kdt).itemsize)
    kb = np.dtype(kdt).itemsize * 8
    for (b0, b1) in [(0, kb), (0, 8), (3, 21), (kb - 13, kb), (0, 1)]:
        keys = rng.integers(0, 2 ** kb, size=n, dtype=kdt)
        if b1 - b0 > 16:  # force ties so stability is tested
            keys[rng.integers(0, n, size=n // 3)] = keys[0]
        vals = None if vdt is None else np.arange(n, dtype=vdt)
        mask = np.uint64(((1 << (b1 - b0)) - 1))
        field = (keys.astype(np.uint64) >> np.uint64(b0)) & mask
        order = np.argsort(field, kind="stable")
        k2 = keys.copy()
        v2 = None if vals is None else vals.copy()
        eng.sort_pairs_host(k2, v2, b0, b1)
        if vals is not None:
            assert (v2 == vals[order]).all(), (n, b0, b1)
            assert (k2 == keys[order]).all(), (n, b0, b1)
        else:
            assert (((k2.astype(np.uint64) >> np.uint64(b0)) & mask) == field[order]).all()
            assert (np.sort(k2) == np.sort(keys)).all()


def test_sort_pairs_skewed_digits(eng):
    n = 300001
    keys = np.zeros(n, np.uint64)
    keys[::7] = 1 << 40
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k2, v2 = keys.copy(), vals.copy()
    eng.sort_pairs_host(k2, v2, 0, 48)
    assert (v2 == vals[order]).all() and (k2 == keys[order]).all()


# ------------------------------------------------------------------------------------------- alphabet (a2)
def test_alphabet_matches_oracle(eng):
    for t in (b"mississippi", G.random_dna(10007, 3), np.arange(256, dtype=np.uint8).repeat(3), G.random_bytes(70001, 9)):
        lut, sigma, bpc = eng.alphabet(t)
        elut, esigma, ebpc = O.alphabet(t)
        assert (lut == elut).all() and sigma == esigma and bpc == ebpc


# ------------------------------------------------------------------------------------------- construct
def _check(eng, text, index_bytes, want_lcp, k=0, exp=None):
    text = np.ascontiguousarray(np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text, np.uint8)
    r = eng.construct(text, index_bytes, want_lcp, k)
    if exp is None:
        exp = O.construct(text, index_bytes * 8, 0, want_lcp)
        assert exp["rc"] == 0 or text.size == 1  # n == 1: the oracle's doubling loop never runs (rc 1), SA = {0}
    assert (r["sa"].astype(np.uint64) == exp["sa"].astype(np.uint64)).all()
    assert (r["isa"].astype(np.uint64) == exp["isa"].astype(np.uint64)).all()
    if want_lcp:
        if text.size == 2:
            # n == 2: the reference's k == 1 raw-character path (kmer.hpp:196-199) makes lcp_bitwise return garbage for
            # LCP[1] (e.g. 2147483647); the engine returns the true LCP -- compare with Kasai (lcp.hpp:46-77) instead
            exp = dict(exp, lcp=O.lcp_from_sa(text, exp["sa"], exp["isa"]))
        assert (r["lcp"].astype(np.uint64) == exp["lcp"].astype(np.uint64)).all()
    return r


def test_mississippi_golden(eng):
    # test/test_psac.cpp:105
    for ib in (4, 8):
        r = eng.construct(b"mississippi", ib, True)
        assert r["sa"].tolist() == [10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2]
        assert r["lcp"].tolist() == [0, 1, 1, 4, 0, 0, 1, 0, 2, 1, 3]
        assert (r["sa"][r["isa"].astype(np.int64)] == np.arange(11)).all()


def test_golden_fixtures_from_the_reference(eng, golden_dir):
    seen = 0
    for f in sorted(glob.glob(os.path.join(golden_dir, "*.npz"))):
        d = np.load(f)
        if "sa" not in d.files or "text" not in d.files:  # (gsa_*.npz: string sets, tests/test_gsa.py)
            continue
        want_lcp = "lcp" in d.files
        exp = dict(sa=d["sa"], isa=d["isa"], lcp=d["lcp"] if want_lcp else None)
        for k in {0, int(d["k"])}:
            _check(eng, d["text"], int(d["index_bytes"]), want_lcp, k, exp)
        seen += 1
    assert seen >= 10


@pytest.mark.parametrize("n", [1, 2, 3, 9, 31, 64, 65, 1000, 1024, 1025, 4097, 66763, 130370])
def test_random_dna_vs_oracle(eng, n):
    t = G.random_dna(n, 100 + n)
    _check(eng, t, 4, False)
    _check(eng, t, 8, True)


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_forced_short_first_key_many_rounds(eng, k):
    # the reference forces more doubling rounds / bucket chasing with small k (test_psac.cpp:148-171)
    t = G.random_dna(20011, 5)
    exp = O.construct(t, 64, 0, True)
    r = _check(eng, t, 8, True, k, exp)
    assert eng.stats()["rounds"] > 1


def test_repetitive_and_periodic_texts(eng):
    for t in (G.repeats_text(3000, 2), G.periodic_text(b"abc", 14681), G.periodic_text(b"a", 5000), G.periodic_text(b"ab", 4097),
              np.zeros(1000, np.uint8)):
        _check(eng, t, 8, True)
        _check(eng, t, 4, True)


def test_bytes_alphabets(eng):
    t255 = np.minimum(G.random_bytes(50000, 4), 254).astype(np.uint8)
    _check(eng, t255, 8, True)
    t256 = G.random_bytes_config4(60000, 4)  # all 256 values: the reference's LUT wraps 0xFF to code 0
    assert len(np.unique(t256)) == 256
    _check(eng, t256, 8, False)
    _check(eng, t256, 4, True)
    _check(eng, G.random_bytes(777, 1) % 20 + 65, 8, True)  # protein-sized alphabet -> 8-bit packing
    _check(eng, G.random_bytes(777, 1) % 13, 4, True)       # 4-bit packing, contains byte 0


def test_user_alphabet_overload(eng):
    t = G.random_dna(5000, 77)
    lut, _, _ = O.alphabet(t)
    r = eng.construct(t, 8, True, lut=lut)
    exp = O.construct(t, 64, 0, True)
    assert (r["sa"] == exp["sa"]).all() and (r["lcp"] == exp["lcp"]).all()


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libpsacref.so not built")
def test_against_unmodified_reference(eng):
    for n, ib, lcp in ((1 << 16, 4, False), (200003, 8, True)):
        t = G.random_dna(n, 31 + n)
        exp = O.ref_construct(t, ib, lcp)
        _check(eng, t, ib, lcp, 0, exp)
    t = G.random_bytes_config4(1 << 16, 8)
    _check(eng, t, 8, False, 0, O.ref_construct(t, 8, False))


def test_medium_size_properties(eng):
    # 16 Mi random DNA (the size of the survey's CPU probe): exact ISA inverse + oracle order check
    n = 1 << 24
    t = G.random_dna(n, 2)
    r = eng.construct(t, 4, True)
    sa, isa, lcp = r["sa"], r["isa"], r["lcp"]
    assert (sa[isa] == np.arange(n, dtype=np.uint32)).all()
    assert O.check_sa(t, sa.astype(np.uint64), isa.astype(np.uint64)) == 0
    # LCP spot check on a sample with the Kasai oracle restricted to a prefix of SA positions
    idx = np.random.default_rng(1).integers(1, n, size=2000)
    for p in idx:
        a, b = int(sa[p - 1]), int(sa[p])
        l = 0
        while a + l < n and b + l < n and t[a + l] == t[b + l]:
            l += 1
        assert l == int(lcp[p])
    assert lcp[0] == 0


def test_sa_class_mirror(eng):
    s = api.SuffixArray(index_bytes=8, construct_lcp=True, engine=eng).construct(b"mississippi")
    assert s.n == 11 and s.local_size == 11 and s.p == 1
    assert s.local_SA.tolist() == [10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2]
    s2 = api.SuffixArray(index_bytes=4, engine=eng).construct_arr(G.random_dna(5000, 1), L=3)
    exp = O.construct(G.random_dna(5000, 1), 32, 0, False)
    assert (s2.local_SA == exp["sa"]).all() and (s2.local_B == exp["isa"]).all() and s2.local_LCP is None


def test_errors(eng):
    with pytest.raises(api.PsacError):
        eng.construct(b"abc", 3)
    r = eng.construct(b"", 8, True)
    assert r["sa"].size == 0


# ------------------------------------------------------------------------------------------- device outputs wider than the internal index
@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 65537, (1 << 20) + 3])
@pytest.mark.parametrize("lcp", [False, True])
def test_device_outputs_64bit_hold_the_32bit_arrays_and_are_widened_in_place(eng, n, lcp):
    """64-bit caller buffers on the device double as storage of the engine's 32-bit arrays and are widened in place
    (engine.cu emit): every size parity, with and without LCP."""
    import torch
    t = G.random_dna(n, 900 + n)
    dev = torch.device("cuda", 0)
    d_t = torch.from_numpy(t).to(dev)
    out = [torch.full((n,), -1, dtype=torch.int64, device=dev) for _ in range(3)]
    torch.cuda.synchronize()
    eng.construct_ptr(d_t.data_ptr(), n, 8, api.LCP if lcp else 0, 0, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr() if lcp else None, device=True)
    exp = O.construct(t, 64, 0, lcp)
    assert (out[0].cpu().numpy().view(np.uint64) == exp["sa"]).all()
    assert (out[1].cpu().numpy().view(np.uint64) == exp["isa"]).all()
    if lcp and n > 2:
        assert (out[2].cpu().numpy().view(np.uint64) == exp["lcp"]).all()


def test_all_256_byte_values_without_lcp_take_the_lean_path(eng):
    """BASELINE configs[3] shape at test size: every byte value occurs (reference LUT overflow: 0xFF gets code 0 = the
    padding), SA / ISA only.  The lean 32-bit-key path must reproduce the reference's order exactly; with LCP the generic
    path is taken and must agree as well."""
    import torch
    t = G.random_bytes_config4((1 << 19) + 5, 41)
    assert len(np.unique(t)) == 256
    exp = O.construct(t, 64, 0, True)
    r = eng.construct(t, 8, False)
    assert (r["sa"] == exp["sa"]).all() and (r["isa"] == exp["isa"]).all()
    assert eng.stats()["sort_elt_bytes"] == 8  # 32-bit carried key + 32-bit suffix index
    r = eng.construct(t, 8, True)
    assert (r["sa"] == exp["sa"]).all() and (r["lcp"] == exp["lcp"]).all()
    assert eng.stats()["sort_elt_bytes"] == 12  # generic path: 64-bit keys
    # device buffers, 64-bit index, no LCP: the bench path of configs[3]
    dev = torch.device("cuda", 0)
    d_t = torch.from_numpy(t).to(dev)
    d_sa = torch.empty(t.size, dtype=torch.int64, device=dev)
    d_isa = torch.empty_like(d_sa)
    torch.cuda.synchronize()
    eng.construct_ptr(d_t.data_ptr(), t.size, 8, 0, 0, d_sa.data_ptr(), d_isa.data_ptr(), None, device=True)
    assert (d_sa.cpu().numpy().view(np.uint64) == exp["sa"]).all() and (d_isa.cpu().numpy().view(np.uint64) == exp["isa"]).all()
    assert eng.check_device_ptr(d_t.data_ptr(), t.size, 8, d_sa.data_ptr(), d_isa.data_ptr(), None)["ok"]


# ------------------------------------------------------------------------------------------- left-branching characters (Lc)
def test_left_branching_chars_match_oracle_and_reference(eng):
    for t in (G.random_dna(100003, 3), G.repeats_text(3000, 2), np.frombuffer(b"mississippi", np.uint8), G.periodic_text(b"abc", 500),
              (G.random_bytes(50000, 8) % 20 + 65).astype(np.uint8), np.frombuffer(b"a", np.uint8)):
        for ib in (4, 8):
            sa = api.SuffixArray(ib, construct_lc=True, engine=eng).construct(t)
            want = O.lc_from_sa_lcp(t, sa.local_SA, sa.local_LCP)
            assert (sa.local_Lc == want).all()
            if O.have_ref() and t.size > 2:
                assert (sa.local_Lc == O.ref_lc(t)).all()


# ------------------------------------------------------------------------------------------- the partitioned SA -> ISA step (n >= 2^23)
def _device_case(eng, t, index_bytes, lcp, k=0):
    """device buffers in and out (used in place by the engine); certified by the device checker, i.e. the reference's
    d_check_sa conditions + LCP by direct comparison over all positions -- the certificate is exact, so a pass means the
    arrays are THE suffix / inverse / LCP arrays"""
    import torch
    dev = torch.device("cuda", 0)
    n = t.size
    d_t = torch.from_numpy(t).to(dev)
    tdt = torch.int64 if index_bytes == 8 else torch.int32
    out = [torch.full((n,), -1, dtype=tdt, device=dev) for _ in range(3)]
    torch.cuda.synchronize()
    eng.construct_ptr(d_t.data_ptr(), n, index_bytes, api.LCP if lcp else 0, k, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr() if lcp else None,
                      device=True)
    st = eng.stats()
    chk = eng.check_device_ptr(d_t.data_ptr(), n, index_bytes, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr() if lcp else None)
    assert chk["ok"], chk
    return st, [o.cpu().numpy().view(np.uint64 if index_bytes == 8 else np.uint32) for o in out]


@pytest.mark.parametrize("index_bytes", [4, 8])
def test_partitioned_isa_step_scatters_positions_and_fixes_up_the_unresolved(eng, index_bytes):
    """n >= 2^23: the SA -> ISA step partitions (suffix, POSITION) pairs generated on the fly and the unresolved suffixes get
    their bucket head afterwards; a 64-bit caller's LCP is written 64 bits wide by the heads kernel.  k = 7: ~3 % of the
    suffixes stay unresolved after the first sort (inside the list capacity); k = 5: most do (the list overflows and the
    bucket-id array is produced after all); k = 0: a handful."""
    n = (1 << 23) + 5
    t = G.random_dna(n, 41)
    for k, lo, hi in ((7, n // 200, n // 16), (5, n // 4, n), (0, 1, 4096)):
        st, out = _device_case(eng, t, index_bytes, True, k)
        assert lo <= st["unresolved_after_first"] <= hi, (k, st["unresolved_after_first"])
        if k == 7:  # element-wise against the CPU oracle once
            exp = O.construct(t, 64, 0, True)
            assert (out[0] == exp["sa"]).all() and (out[1] == exp["isa"]).all() and (out[2] == exp["lcp"]).all()


def test_partitioned_isa_step_on_repetitive_and_periodic_texts(eng):
    for t in (G.periodic_text(b"abc", (1 << 23) // 3 + 11), G.repeats_text(1 << 21, 5)[: (1 << 23) + 1]):
        for ib, lcp in ((8, True), (4, False)):
            st, _ = _device_case(eng, t, ib, lcp)
            assert st["rounds"] > 2
