"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front end of the two CPU oracles.

* ``liboracle.so``      : oracle/psac_oracle.c, the plain-C restatement (kind "port").
* ``_ref/libpsacref.so``: the UNMODIFIED reference compiled at np=1 against the MPI shim
                          (oracle/ref_driver.cpp, kind "reference").  Built only where
                          /root/reference exists; the prebuilt file travels to the GPU box.
* ``_ref/libdivsufsort64.so``: the reference's vendored libdivsufsort (second, independent oracle for SA:
                          induced sorting, orders by unsigned byte).

All arrays are numpy; indices are uint64 in the port, uint32/uint64 (``index_bytes``) in the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpsacref.so")
DSS_SO = os.path.join(HERE, "_ref", "libdivsufsort64.so")

_port = None
_ref = None


def build(verbose=False):
    """Compile liboracle.so (always) and _ref/libpsacref.so (only where /root/reference exists)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile"), PORT_SO], stdout=out)
    if os.path.isdir("/root/reference/include"):
        subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile"), REF_SO, DSS_SO], stdout=out)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        _port = C.CDLL(PORT_SO)
    return _port


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libpsacref.so missing (build it where /root/reference exists)")
        _ref = C.CDLL(REF_SO)
        _ref.psacref_init()
    return _ref


def _text(t):
    if isinstance(t, (bytes, bytearray)):
        t = np.frombuffer(bytes(t), dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    return t


# ----------------------------------------------------------------------------- port (plain C)

def alphabet(text):
    t = _text(text)
    lut = np.zeros(256, np.uint8)
    sigma = C.c_uint()
    bpc = C.c_uint()
    port().oracle_alphabet(_vp(t), C.c_size_t(t.size), _vp(lut), C.byref(sigma), C.byref(bpc))
    return lut, sigma.value, bpc.value


def optimal_k(bits_per_char, index_bits, min_local_size, p=1, k=0):
    f = port().oracle_optimal_k
    f.restype = C.c_uint
    return f(C.c_uint(bits_per_char), C.c_uint(index_bits), C.c_size_t(min_local_size), C.c_int(p), C.c_uint(k))


def kmer_generation(text, lut, l, k, index_bits):
    t = _text(text)
    out = np.zeros(t.size, np.uint64)
    port().oracle_kmer_generation(_vp(t), C.c_size_t(t.size), _vp(np.ascontiguousarray(lut, np.uint8)), C.c_uint(l), C.c_uint(k),
                                  C.c_uint(index_bits), _vp(out))
    return out


def shift(b, h):
    b = np.ascontiguousarray(b, np.uint64)
    out = np.zeros_like(b)
    port().oracle_shift(_vp(b), C.c_size_t(b.size), C.c_size_t(h), _vp(out))
    return out


def idxsort(v1, v2):
    v1 = np.array(v1, np.uint64)
    v2 = np.array(v2, np.uint64)
    idx = np.zeros(v1.size, np.uint64)
    rc = port().oracle_idxsort(_vp(v1), _vp(v2), C.c_size_t(v1.size), _vp(idx))
    assert rc == 0
    return v1, v2, idx


def rebucket(v1, v2):
    v1 = np.array(v1, np.uint64)
    v2 = np.ascontiguousarray(v2, np.uint64)
    ub = C.c_uint64()
    ue = C.c_uint64()
    port().oracle_rebucket(_vp(v1), _vp(v2), C.c_size_t(v1.size), C.byref(ub), C.byref(ue))
    return v1, ub.value, ue.value


def bulk_permute(vec, idx):
    vec = np.ascontiguousarray(vec, np.uint64)
    idx = np.ascontiguousarray(idx, np.uint64)
    out = np.zeros_like(vec)
    port().oracle_bulk_permute(_vp(vec), _vp(idx), C.c_size_t(vec.size), _vp(out))
    return out


def lcp_bitwise(x, y, k, l, word_bits):
    f = port().oracle_lcp_bitwise
    f.restype = C.c_uint
    return f(C.c_uint64(x), C.c_uint64(y), C.c_uint(k), C.c_uint(l), C.c_uint(word_bits))


def initial_kmer_lcp(b1, b2, k, l, word_bits):
    b1 = np.ascontiguousarray(b1, np.uint64)
    b2 = np.ascontiguousarray(b2, np.uint64)
    lcp = np.zeros(b1.size, np.uint64)
    port().oracle_initial_kmer_lcp(_vp(b1), _vp(b2), C.c_size_t(b1.size), C.c_uint(k), C.c_uint(l), C.c_uint(word_bits), _vp(lcp))
    return lcp


def resolve_next_lcp(b1, b2, dist, lcp):
    b1 = np.ascontiguousarray(b1, np.uint64)
    b2 = np.ascontiguousarray(b2, np.uint64)
    lcp = np.array(lcp, np.uint64)
    rc = port().oracle_resolve_next_lcp(_vp(b1), _vp(b2), C.c_size_t(b1.size), C.c_uint64(dist), _vp(lcp))
    assert rc == 0
    return lcp


def construct(text, index_bits=64, k=0, want_lcp=False):
    """Port of suffix_array::construct at np=1.  Returns dict(sa, isa, lcp|None, rounds, rc)."""
    t = _text(text)
    n = t.size
    sa = np.zeros(n, np.uint64)
    isa = np.zeros(n, np.uint64)
    lcp = np.zeros(n, np.uint64) if want_lcp else None
    rounds = C.c_uint()
    rc = port().oracle_construct(_vp(t), C.c_size_t(n), C.c_uint(index_bits), C.c_uint(k), _vp(sa), _vp(isa), _vp(lcp), C.byref(rounds))
    if rc < 0:
        raise MemoryError("oracle_construct failed")
    return dict(sa=sa, isa=isa, lcp=lcp, rounds=rounds.value, rc=rc)


def stringset_alphabet(flat, sep=ord("$")):
    """lut / sigma / bits_per_char of alphabet<char>::from_string over the non-separator characters (test/test_gsa.cpp:86)"""
    t = _text(flat)
    return alphabet(t[t != sep])


def construct_ss(flat, sep=ord("$"), index_bits=64, want_lcp=True, lut=None, bits_per_char=None):
    """Port of suffix_array::construct_ss (generalized SA of the `sep`-separated strings) at np=1.
    Positions index the concatenation without separators.  Returns dict(sa, isa, lcp|None, rounds, n)."""
    t = _text(flat)
    if lut is None:
        lut, _, bits_per_char = stringset_alphabet(t, sep)
    cap = max(t.size, 1)
    sa = np.zeros(cap, np.uint64)
    isa = np.zeros(cap, np.uint64)
    lcp = np.zeros(cap, np.uint64) if want_lcp else None
    rounds = C.c_uint()
    f = port().oracle_construct_ss
    f.restype = C.c_long
    m = f(_vp(t), C.c_size_t(t.size), C.c_ubyte(sep), _vp(np.ascontiguousarray(lut, np.uint8)), C.c_uint(bits_per_char), C.c_uint(index_bits),
          _vp(sa), _vp(isa), _vp(lcp), C.byref(rounds))
    if m < 0:
        raise RuntimeError("oracle_construct_ss failed: %d" % m)
    return dict(sa=sa[:m], isa=isa[:m], lcp=None if lcp is None else lcp[:m], rounds=rounds.value, n=m)


def gsa_naive(flat, sep=ord("$")):
    """The definition the generalized SA must satisfy (test/test_gsa.cpp:97-98 and the closed forms :31-66): suffixes of
    every string compared as strings (a proper prefix is smaller), ties by position; LCP = common prefix length."""
    t = bytes(_text(flat))
    suf = []
    pos = 0
    for s in t.split(bytes([sep])):
        for i in range(len(s)):
            suf.append((s[i:], pos + i))
        pos += len(s)
    suf.sort()
    m = len(suf)
    sa = np.array([p for _, p in suf], np.uint64).reshape(m)
    lcp = np.zeros(m, np.uint64)
    for q in range(1, m):
        a, b = suf[q - 1][0], suf[q][0]
        l = 0
        while l < len(a) and l < len(b) and a[l] == b[l]:
            l += 1
        lcp[q] = l
    isa = np.zeros(m, np.uint64)
    isa[sa.astype(np.int64)] = np.arange(m, dtype=np.uint64)
    return dict(sa=sa, isa=isa, lcp=lcp, n=m)


def construct_wide(text, index_bits=64, want_lcp=False):
    """Texts over wide characters (reference suffix_array<int, ...> with int_alphabet, include/alphabet.hpp:355-513: the
    characters are ordered by VALUE).  SA / ISA / LCP depend only on the order and equality of the characters, so the port
    ranks the distinct values (<= 255 of them) into bytes and runs `construct`; pinned to the unmodified reference in
    tests/test_wide.py."""
    t = np.ascontiguousarray(text)
    vals, inv = np.unique(t, return_inverse=True)
    if vals.size > 255:
        raise ValueError("more than 255 distinct characters")
    return construct((inv + 1).astype(np.uint8), index_bits, 0, want_lcp)


def construct_arr(text, L, index_bits=64):
    t = _text(text)
    n = t.size
    sa = np.zeros(n, np.uint64)
    isa = np.zeros(n, np.uint64)
    rounds = C.c_uint()
    rc = port().oracle_construct_arr(_vp(t), C.c_size_t(n), C.c_uint(index_bits), C.c_int(L), _vp(sa), _vp(isa), C.byref(rounds))
    if rc < 0:
        raise RuntimeError("oracle_construct_arr failed: %d" % rc)
    return dict(sa=sa, isa=isa, rounds=rounds.value, rc=rc)


def lcp_from_sa(text, sa, isa):
    t = _text(text)
    sa = np.ascontiguousarray(sa, np.uint64)
    isa = np.ascontiguousarray(isa, np.uint64)
    lcp = np.zeros(t.size, np.uint64)
    port().oracle_lcp_from_sa(_vp(t), C.c_size_t(t.size), _vp(sa), _vp(isa), _vp(lcp))
    return lcp


def sa_naive(text):
    t = _text(text)
    sa = np.zeros(t.size, np.uint64)
    port().oracle_sa_naive(_vp(t), C.c_size_t(t.size), _vp(sa))
    return sa


def check_sa(text, sa, isa, lut=None):
    t = _text(text)
    if lut is None:
        lut = alphabet(t)[0]
    sa = np.ascontiguousarray(sa, np.uint64)
    isa = np.ascontiguousarray(isa, np.uint64)
    return port().oracle_check_sa(_vp(t), C.c_size_t(t.size), _vp(np.ascontiguousarray(lut, np.uint8)), _vp(sa), _vp(isa))


def ansv_sequential(vals, left, nonsv=0):
    v = np.ascontiguousarray(vals, np.uint64)
    out = np.zeros(v.size, np.uint64)
    rc = port().oracle_ansv_sequential(_vp(v), C.c_size_t(v.size), C.c_int(1 if left else 0), C.c_uint64(nonsv), _vp(out))
    assert rc == 0
    return out


def ansv(vals, left_type=0, right_type=0, nonsv=0):
    v = np.ascontiguousarray(vals, np.uint64)
    l = np.zeros(v.size, np.uint64)
    r = np.zeros(v.size, np.uint64)
    port().oracle_ansv(_vp(v), C.c_size_t(v.size), C.c_int(left_type), C.c_int(right_type), C.c_uint64(nonsv), _vp(l), _vp(r))
    return l, r


# ------------------------------------------------------------- reference (unmodified, np=1)

def ref_construct(text, index_bytes=8, want_lcp=False, k=0, fast=True, arr_L=0):
    t = _text(text)
    n = t.size
    dt = np.uint32 if index_bytes == 4 else np.uint64
    sa = np.zeros(n, dt)
    isa = np.zeros(n, dt)
    lcp = np.zeros(n, dt) if want_lcp else None
    rc = ref().psacref_construct(_vp(t), C.c_size_t(n), C.c_int(index_bytes), C.c_int(1 if want_lcp else 0), C.c_uint(k),
                                 C.c_int(1 if fast else 0), C.c_int(arr_L), _vp(sa), _vp(isa), _vp(lcp))
    if rc != 0:
        raise RuntimeError("psacref_construct rc=%d" % rc)
    return dict(sa=sa, isa=isa, lcp=lcp)


def ref_alphabet(text):
    t = _text(text)
    lut = np.zeros(256, np.uint8)
    sigma = C.c_uint()
    bpc = C.c_uint()
    ref().psacref_alphabet(_vp(t), C.c_size_t(t.size), _vp(lut), C.byref(sigma), C.byref(bpc))
    return lut, sigma.value, bpc.value


def ref_optimal_k(text, index_bytes, k=0):
    t = _text(text)
    f = ref().psacref_optimal_k
    f.restype = C.c_uint
    return f(_vp(t), C.c_size_t(t.size), C.c_int(index_bytes), C.c_uint(k))


def ref_kmer_generation(text, index_bytes, k):
    t = _text(text)
    out = np.zeros(t.size, np.uint32 if index_bytes == 4 else np.uint64)
    ref().psacref_kmer_generation(_vp(t), C.c_size_t(t.size), C.c_int(index_bytes), C.c_uint(k), _vp(out))
    return out


def ref_lcp_bitwise(x, y, k, l, word_bits):
    if word_bits == 32:
        f = ref().psacref_lcp_bitwise32
        f.restype = C.c_uint
        return f(C.c_uint32(x), C.c_uint32(y), C.c_uint(k), C.c_uint(l))
    f = ref().psacref_lcp_bitwise64
    f.restype = C.c_uint
    return f(C.c_uint64(x), C.c_uint64(y), C.c_uint(k), C.c_uint(l))


def ref_lcp_from_sa(text, sa, isa):
    t = _text(text)
    sa = np.ascontiguousarray(sa, np.uint64)
    isa = np.ascontiguousarray(isa, np.uint64)
    lcp = np.zeros(t.size, np.uint64)
    ref().psacref_lcp_from_sa(_vp(t), C.c_size_t(t.size), _vp(sa), _vp(isa), _vp(lcp))
    return lcp


def ref_ansv_sequential(vals, left, nonsv=0):
    v = np.ascontiguousarray(vals, np.uint64)
    out = np.zeros(v.size, np.uint64)
    ref().psacref_ansv_sequential(_vp(v), C.c_size_t(v.size), C.c_int(1 if left else 0), C.c_uint64(nonsv), _vp(out))
    return out


def ref_ansv(vals, left_type=0, right_type=0, nonsv=0):
    v = np.ascontiguousarray(vals, np.uint64)
    l = np.zeros(v.size, np.uint64)
    r = np.zeros(v.size, np.uint64)
    rc = ref().psacref_ansv(_vp(v), C.c_size_t(v.size), C.c_int(left_type), C.c_int(right_type), C.c_uint64(nonsv), _vp(l), _vp(r))
    if rc != 0:
        raise RuntimeError("psacref_ansv rc=%d" % rc)
    return l, r


def ref_suffix_tree(text):
    t = _text(text)
    n = t.size
    cap = 257 * n
    nodes = np.zeros(cap, np.uint64)
    f = ref().psacref_suffix_tree
    f.restype = C.c_long
    w = f(_vp(t), C.c_size_t(n), _vp(nodes), C.c_size_t(cap))
    if w < 0:
        raise RuntimeError("psacref_suffix_tree rc=%d" % w)
    return nodes[: w * n].reshape(n, w).copy()


def ref_write(text, index_bytes, basename):
    """the UNMODIFIED reference builds SA + LCP of `text` and writes <basename>.sa / .lcp / .alpha with its own write()"""
    t = _text(text)
    rc = ref().psacref_write(_vp(t), C.c_size_t(t.size), C.c_int(index_bytes), str(basename).encode())
    if rc != 0:
        raise RuntimeError("psacref_write rc=%d" % rc)


def ref_read(basename, index_bytes, cap):
    """the UNMODIFIED reference reads <basename>.sa / .lcp / .alpha with its own read(): dict(n, sa, lcp, lut, sigma)"""
    dt = np.uint32 if index_bytes == 4 else np.uint64
    sa = np.zeros(cap, dt)
    lcp = np.zeros(cap, dt)
    lut = np.zeros(256, np.uint8)
    sigma = C.c_uint()
    f = ref().psacref_read
    f.restype = C.c_long
    n = f(str(basename).encode(), C.c_int(index_bytes), _vp(sa), _vp(lcp), C.c_size_t(cap), _vp(lut), C.byref(sigma))
    if n < 0:
        raise RuntimeError("psacref_read rc=%d" % n)
    return dict(n=int(n), sa=sa[:n], lcp=lcp[:n], lut=lut, sigma=sigma.value)


def lc_from_sa_lcp(text, sa, lcp):
    """left-branching characters restated: Lc[i] = text[SA[i-1] + LCP[i]], 0 past the end and for i = 0
    (reference include/desa.hpp:296-312 recomputes local_Lc exactly like this; suffix_array.hpp:1365-1383 by k-mer decoding)"""
    t = _text(text)
    n = t.size
    out = np.zeros(n, np.uint8)
    if n > 1:
        g = np.asarray(sa[:-1]).astype(np.int64) + np.asarray(lcp[1:]).astype(np.int64)
        ok = g < n
        out[1:][ok] = t[g[ok]]
    return out


def ref_lc(text, k=0):
    """local_Lc of the UNMODIFIED reference built with _CONSTRUCT_LC (oracle/_ref)"""
    t = _text(text)
    out = np.zeros(t.size, np.uint8)
    rc = ref().psacref_lc(_vp(t), C.c_size_t(t.size), C.c_uint(k), _vp(out))
    if rc != 0:
        raise RuntimeError("psacref_lc rc=%d" % rc)
    return out


def ref_construct_ss(flat, sep=ord("$"), index_bytes=8, alpha_chars=None):
    """suffix_array<char,index_t,true>::construct_ss of the UNMODIFIED reference at np=1 (oracle/_ref)"""
    t = _text(flat)
    if alpha_chars is None:
        alpha_chars = bytes(np.unique(t[t != sep]).tolist())
    dt = np.uint32 if index_bytes == 4 else np.uint64
    cap = max(t.size, 1)
    sa = np.zeros(cap, dt)
    isa = np.zeros(cap, dt)
    lcp = np.zeros(cap, dt)
    f = ref().psacref_construct_ss
    f.restype = C.c_long
    m = f(_vp(t), C.c_size_t(t.size), C.c_char(bytes([sep])), alpha_chars, C.c_size_t(len(alpha_chars)), C.c_int(index_bytes), _vp(sa), _vp(isa),
          _vp(lcp), C.c_size_t(cap))
    if m < 0:
        raise RuntimeError("psacref_construct_ss rc=%d" % m)
    return dict(sa=sa[:m], isa=isa[:m], lcp=lcp[:m], n=m)


def ref_construct_int(text, index_bytes=8, want_lcp=False):
    """suffix_array<int, index_t, LCP>::construct of the UNMODIFIED reference at np=1 (int_alphabet path)"""
    t = np.ascontiguousarray(text, np.int32)
    n = t.size
    dt = np.uint32 if index_bytes == 4 else np.uint64
    sa, isa = np.zeros(n, dt), np.zeros(n, dt)
    lcp = np.zeros(n, dt) if want_lcp else None
    rc = ref().psacref_construct_int(_vp(t), C.c_size_t(n), C.c_int(index_bytes), C.c_int(1 if want_lcp else 0), _vp(sa), _vp(isa), _vp(lcp))
    if rc != 0:
        raise RuntimeError("psacref_construct_int rc=%d" % rc)
    return dict(sa=sa, isa=isa, lcp=lcp)


def ref_rand_dna(n, seed):
    out = np.zeros(n, np.uint8)
    ref().psacref_rand_dna(C.c_size_t(n), C.c_int(seed), _vp(out))
    return out


# ------------------------------------------------------------- libdivsufsort (second, independent oracle)
_dss = None


def have_dss():
    return os.path.exists(DSS_SO)


def dss_sa(text):
    """Suffix array by the reference's vendored libdivsufsort (divsufsort64; reference include/divsufsort_wrapper.hpp:54-71).
    Byte order: equals psac's SA when at most 255 distinct byte values occur."""
    global _dss
    if _dss is None:
        if not have_dss():
            raise RuntimeError("oracle/_ref/libdivsufsort64.so missing (build it where /root/reference exists)")
        _dss = C.CDLL(DSS_SO)
    t = _text(text)
    sa = np.zeros(t.size, np.int64)
    rc = _dss.divsufsort64(_vp(t), _vp(sa), C.c_int64(t.size))
    if rc != 0:
        raise RuntimeError("divsufsort64 rc=%d" % rc)
    return sa.astype(np.uint64)


def dss_check(text, sa):
    """libdivsufsort's own checker (sufcheck64; divsufsort_wrapper.hpp:90-109): 0 = valid suffix array"""
    global _dss
    if _dss is None:
        dss_sa(b"a")
    t = _text(text)
    s = np.ascontiguousarray(sa, np.int64)
    return int(_dss.sufcheck64(_vp(t), _vp(s), C.c_int64(t.size), C.c_int32(0)))
