"""Device-side certificate (psacb200_check*, reference d_check_sa + check_lcp, include/check_suffix_array.hpp:151-267):
accepts the engine's results, and flags every kind of corruption -- including a wrong order with a consistent inverse,
which an ISA[SA[i]] == i test alone cannot see."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from psac_b200 import api, textgen as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


TEXTS = [
    ("dna", lambda: G.random_dna(200003, 5)),
    ("dna tiny", lambda: G.random_dna(37, 6)),
    ("one char", lambda: np.frombuffer(b"a", np.uint8).copy()),
    ("periodic", lambda: G.periodic_text(b"abc", 5000)),
    ("repeats", lambda: G.repeats_text(3000, 9)),
    ("bytes 256 quirk", lambda: G.random_bytes_config4(1 << 17, 7)),
    ("protein", lambda: (G.random_bytes(50000, 8) % 20 + 65).astype(np.uint8)),
]


@pytest.mark.parametrize("name,make", TEXTS, ids=[t[0] for t in TEXTS])
@pytest.mark.parametrize("ib", [4, 8])
def test_check_accepts_engine_and_oracle_results(eng, name, make, ib):
    t = make()
    r = eng.construct(t, ib, True)
    rep = eng.check(t, r["sa"], r["isa"], r["lcp"])
    assert rep["ok"] and rep["checked_lcp"] == 1 and rep["first_bad"] == 2**64 - 1, rep
    exp = O.construct(t, ib * 8, 0, True)
    dt = np.uint32 if ib == 4 else np.uint64
    rep = eng.check(t, exp["sa"].astype(dt), exp["isa"].astype(dt), exp["lcp"].astype(dt))
    assert rep["ok"], rep
    rep = eng.check(t, r["sa"], r["isa"])  # SA / ISA only
    assert rep["ok"] and rep["checked_lcp"] == 0


def test_check_flags_corruptions(eng):
    t = G.random_dna(100000, 21)
    r = eng.construct(t, 8, True)
    sa, isa, lcp = r["sa"], r["isa"], r["lcp"]
    n = t.size
    # out of range
    bad = sa.copy()
    bad[17] = n + 5
    rep = eng.check(t, bad, isa, lcp)
    assert rep["bad_range"] == 1 and not rep["ok"] and rep["first_bad"] <= 17
    # not a permutation (duplicate)
    bad = sa.copy()
    bad[100] = bad[200]
    rep = eng.check(t, bad, isa, lcp)
    assert rep["bad_inverse"] >= 1 and not rep["ok"]
    # ISA entry wrong
    bad = isa.copy()
    bad[int(sa[500])] = 501
    rep = eng.check(t, sa, bad, lcp)
    assert rep["bad_inverse"] >= 1 and not rep["ok"]
    # LCP entry wrong
    bad = lcp.copy()
    bad[777] += 1
    rep = eng.check(t, sa, isa, bad)
    assert rep["bad_lcp"] == 1 and rep["bad_order"] == 0 and rep["bad_inverse"] == 0 and rep["first_bad"] == 777
    bad = lcp.copy()
    bad[0] = 3
    assert eng.check(t, sa, isa, bad)["bad_lcp"] == 1
    # wrong ORDER with a CONSISTENT inverse: two adjacent suffixes exchanged in SA and ISA
    bsa, bisa = sa.copy(), isa.copy()
    i = 4242
    bsa[i], bsa[i + 1] = sa[i + 1], sa[i]
    bisa[int(bsa[i])], bisa[int(bsa[i + 1])] = i, i + 1
    assert (bisa[bsa] == np.arange(n)).all()  # the old certificate accepts this
    rep = eng.check(t, bsa, bisa)
    assert rep["bad_inverse"] == 0 and rep["bad_order"] >= 1 and not rep["ok"]


def test_check_device_pointers(eng):
    import torch
    t = G.random_dna(1 << 20, 31)
    dev = torch.device("cuda", 0)
    d_t = torch.from_numpy(t).to(dev)
    d_sa = torch.empty(t.size, dtype=torch.int64, device=dev)
    d_isa = torch.empty_like(d_sa)
    d_lcp = torch.empty_like(d_sa)
    torch.cuda.synchronize()
    eng.construct_ptr(d_t.data_ptr(), t.size, 8, api.LCP, 0, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr(), device=True)
    rep = eng.check_device_ptr(d_t.data_ptr(), t.size, 8, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr())
    assert rep["ok"] and rep["n"] == t.size
    d_lcp[12345] += 2
    torch.cuda.synchronize()
    rep = eng.check_device_ptr(d_t.data_ptr(), t.size, 8, d_sa.data_ptr(), d_isa.data_ptr(), d_lcp.data_ptr())
    assert rep["bad_lcp"] == 1 and rep["first_bad"] == 12345


SAFE_RANK_SCRIPT = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from oracle import pyoracle as O
from psac_b200 import api, textgen as G
import ctypes as C
e = api.Engine(0)
assert api.lib().psacb200_rank_mode(e._h, None) == 1, "PSACB200_SAFE_RANK=1 must select the match.any ranking"
for t, ib, k in [(G.random_dna(300007, 3), 4, 0), (G.random_dna(70001, 4), 8, 3), (G.repeats_text(4000, 5), 8, 0),
                 (G.random_bytes_config4(1 << 17, 6), 8, 0), ((G.random_bytes(90000, 8) %% 20 + 65).astype(np.uint8), 4, 0)]:
    r = e.construct(t, ib, True, k=k)
    x = O.construct(t, ib * 8, k, True)
    assert (r["sa"] == x["sa"]).all() and (r["isa"] == x["isa"]).all() and (r["lcp"] == x["lcp"]).all()
keys = (G.splitmix64(9, np.arange(1 << 20, dtype=np.uint64)) >> np.uint64(20)).astype(np.uint64)
vals = np.arange(keys.size, dtype=np.uint32)
order = np.argsort(keys, kind="stable")
e.sort_pairs_host(keys, vals, 0, 44)
assert (vals == order.astype(np.uint32)).all()
print("safe-rank ok")
"""


def test_match_any_ranking_path_matches_oracle():
    """The second ranking path (documented warp primitives only) that replaces the single-ATOMS ranking when the hardware
    self-test fails: forced with PSACB200_SAFE_RANK=1 in a fresh process, parity with the oracle and with a stable argsort."""
    env = dict(os.environ, PSACB200_SAFE_RANK="1")
    r = subprocess.run([sys.executable, "-c", SAFE_RANK_SCRIPT % ROOT], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "safe-rank ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
