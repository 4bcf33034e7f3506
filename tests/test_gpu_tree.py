"""GPU parity of ANSV (reference include/ansv.hpp) and suffix-tree construction (include/suffix_tree.hpp) through the C ABI,
against the CPU oracle, the golden fixtures generated from the unmodified reference, and -- where oracle/_ref travels --
the unmodified reference itself.  Mirrors test/test_ansv.cpp:316-326 (sizes 13/137/1000/26666, rand() % 100, all mode
combinations) and test/test_suffixtree.cpp:66-162 (mississippi golden table, random DNA, (abc)^n)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from psac_b200 import api
from psac_b200 import textgen as G

pytestmark = pytest.mark.gpu
NONSV = np.uint64(2**64 - 1)


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("n", [1, 2, 13, 137, 1000, 26666])
@pytest.mark.parametrize("dt", [np.uint32, np.uint64])
def test_ansv_all_mode_combinations_vs_oracle(eng, n, dt):
    rng = np.random.default_rng(n)
    for vals in (rng.integers(0, 100, size=n), rng.integers(0, 3, size=n), np.arange(n), np.arange(n)[::-1].copy(), np.zeros(n, np.int64),
                 rng.permutation(n)):
        v = vals.astype(dt)
        if n > 5000 and vals is not None and len(np.unique(vals)) < 4:
            continue  # the oracle is O(n^2) on long runs of equal values
        for lt in (0, 1, 2):
            for rt in (0, 1, 2):
                l, r = eng.ansv(v, lt, rt, int(NONSV))
                el, er = O.ansv(v.astype(np.uint64), lt, rt, int(NONSV))
                assert (l == el).all(), (n, lt, rt)
                assert (r == er).all(), (n, lt, rt)


def test_ansv_golden_from_the_reference(eng, golden_dir):
    d = np.load(os.path.join(golden_dir, "ansv137.npz"))
    vals = d["vals"]
    nonsv = 2 ** 64 - 1
    for lt in range(3):
        for rt in range(3):
            l, r = eng.ansv(vals.astype(np.uint64), lt, rt, nonsv)
            assert (l == d["l_%d_%d" % (lt, rt)]).all() and (r == d["r_%d_%d" % (lt, rt)]).all()


def test_ansv_large_against_sequential_oracle(eng):
    # nearest_sm both sides at a size where only the O(n) stack oracle (ansv.hpp:47-65) is practical
    n = 3_000_017
    v = np.random.default_rng(3).integers(0, 40, size=n).astype(np.uint32)
    l, r = eng.ansv(v, 0, 0, int(NONSV))
    assert (l == O.ansv_sequential(v, True, int(NONSV))).all()
    assert (r == O.ansv_sequential(v, False, int(NONSV))).all()


def test_suffix_tree_mississippi_golden(eng, golden_dir):
    d = np.load(os.path.join(golden_dir, "stree_mississippi.npz"))
    text = np.frombuffer(b"mississippi", np.uint8)
    exp = O.construct(text, 64, 0, True)
    nodes = eng.suffix_tree(text, exp["sa"], exp["lcp"])
    assert nodes.shape == (11, 5)
    assert (nodes.reshape(-1) == d["nodes"].reshape(-1)).all()  # test_suffixtree.cpp:66-79


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libpsacref.so not built")
@pytest.mark.parametrize("ib", [4, 8])
def test_suffix_tree_vs_unmodified_reference(eng, ib):
    dt = np.uint32 if ib == 4 else np.uint64
    texts = [G.random_dna(n, 13 + n) for n in (116, 1000, 23713)] + [G.periodic_text(b"abc", k) for k in (3, 25, 97, 151)] + [
        np.frombuffer(b"mississippi", np.uint8), G.random_bytes(5000, 2) % 20 + 65, G.repeats_text(300, 5)]
    for t in texts:
        t = np.ascontiguousarray(t, np.uint8)
        exp = O.construct(t, 64, 0, True)
        got = eng.suffix_tree(t, exp["sa"].astype(dt), exp["lcp"].astype(dt))
        ref = O.ref_suffix_tree(t)
        assert got.shape == ref.shape and (got == ref).all(), t.size


def test_suffix_tree_end_to_end_from_engine_outputs(eng):
    # construct() on the GPU, then the tree from ITS outputs: every leaf n + i hangs below exactly one parent and the
    # number of occupied cells is n leaves + (number of distinct internal nodes other than the root)
    t = G.random_dna(177861, 13)  # test_suffixtree.cpp:101
    r = eng.construct(t, 8, True)
    nodes = eng.suffix_tree(t, r["sa"], r["lcp"])
    n = t.size
    flat = nodes.reshape(-1)
    occ = flat[flat != 0]
    leaves = occ[occ >= n]
    assert leaves.size == n and np.unique(leaves).size == n
    inner = occ[occ < n]
    assert np.unique(inner).size == inner.size
    if O.have_ref():
        assert (nodes == O.ref_suffix_tree(t)).all()


# ------------------------------------------------------------------------------------------- device-resident entry points
def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64 if a.dtype.itemsize == 8 else (np.int32 if a.dtype.itemsize == 4 else np.uint8))).to("cuda:0")


@pytest.mark.parametrize("ib", [4, 8])
def test_device_suffix_tree_fused_matches_reference_table(eng, ib):
    """psacb200_suffix_tree_device (ANSV searched on the fly, device arrays) against the host entry point, which is
    pinned to the unmodified reference above, and against the reference itself where it travels."""
    import torch
    dt = np.uint32 if ib == 4 else np.uint64
    texts = [G.random_dna(n, 31 + n) for n in (1, 2, 116, 23713, 300007)] + [G.periodic_text(b"abc", 151), np.frombuffer(b"mississippi", np.uint8),
                                                                             (G.random_bytes(60000, 2) % 20 + 65).astype(np.uint8), G.repeats_text(2000, 5),
                                                                             G.random_dna(70001, 77)]
    for t in texts:
        t = np.ascontiguousarray(t, np.uint8)
        r = eng.construct(t, ib, True)
        want = eng.suffix_tree(t, r["sa"], r["lcp"])
        d_t, d_sa, d_lcp = _dev(t), _dev(r["sa"].astype(dt)), _dev(r["lcp"].astype(dt))
        width = want.shape[1]
        d_nodes = torch.full((t.size, width), -1, dtype=torch.int64, device="cuda:0")
        torch.cuda.synchronize()
        sigma = eng.suffix_tree_device_ptr(d_t.data_ptr(), t.size, ib, d_sa.data_ptr(), d_lcp.data_ptr(), d_nodes.data_ptr(), d_nodes.numel())
        assert sigma + 1 == width
        got = d_nodes.cpu().numpy().view(np.uint64)
        assert (got == want).all(), (t.size, ib)
        if O.have_ref() and t.size > 2 and ib == 8:  # (n <= 2: the reference's LCP[1] is uninitialised, tests/test_gpu_parity.py)
            assert (got == O.ref_suffix_tree(t)).all()


def test_device_ansv_matches_host_api(eng):
    import torch
    rng = np.random.default_rng(11)
    for n in (1, 33, 1000, 200003):
        for dt in (np.uint32, np.uint64):
            v = rng.integers(0, 50, size=n).astype(dt)
            d_v = _dev(v)
            d_l = torch.empty(n, dtype=torch.int64, device="cuda:0")
            d_r = torch.empty(n, dtype=torch.int64, device="cuda:0")
            torch.cuda.synchronize()
            for lt, rt in ((0, 0), (2, 0), (1, 2)):
                eng.ansv_device_ptr(d_v.data_ptr(), n, v.dtype.itemsize, lt, rt, int(NONSV), d_l.data_ptr(), d_r.data_ptr())
                l, r = eng.ansv(v, lt, rt, int(NONSV))
                assert (d_l.cpu().numpy().view(np.uint64) == l).all() and (d_r.cpu().numpy().view(np.uint64) == r).all(), (n, lt, rt)
