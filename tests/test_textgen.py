"""Synthetic inputs (own counter-based PRNG so that any shard can generate its block): the numpy
generators the parity tests / the oracle use and the torch generators the bench uses on the device must produce the same
text, block by block, and the string-set generator must be deterministic and well formed."""
import numpy as np
import pytest

from psac_b200 import textgen as G


def test_numpy_and_torch_generators_agree():
    torch = pytest.importorskip("torch")
    dev = torch.device("cpu")
    for n, seed in ((1, 1), (1000, 2), (65537, 3), (1 << 18, 5)):
        assert np.array_equal(G.random_dna(n, seed), G.random_dna_torch(n, seed, dev, chunk=1 << 16).numpy())
        assert np.array_equal(G.random_bytes(n, seed), G.random_bytes_torch(n, seed, dev, chunk=1 << 16).numpy())


def test_blocks_of_a_text_can_be_generated_independently():
    n, seed = 100003, 7
    whole = G.random_dna(n, seed)
    for start, size in ((0, 10), (12345, 50000), (n - 7, 7)):
        assert np.array_equal(whole[start:start + size], G.random_dna(size, seed, start))
    wb = G.random_bytes(n, seed)
    assert np.array_equal(wb[999:5000], G.random_bytes(4001, seed, 999))


def test_dna_alphabet_and_byte_coverage():
    t = G.random_dna(1 << 16, 11)
    assert set(np.unique(t).tolist()) == set(b"ACGT")
    b = G.random_bytes_config4(1 << 16, 4)
    assert np.unique(b).size == 256 and b[-1] != 0xFF  # configs[3]: all 256 values, last byte != 0xFF


def test_random_stringset_is_deterministic_and_well_formed():
    a = G.random_stringset(300, 150, 31)
    assert np.array_equal(a, G.random_stringset(300, 150, 31))
    strs = a.tobytes().split(b"$")
    assert len(strs) == 300 and all(1 <= len(s) <= 150 for s in strs)
    assert set(np.unique(a).tolist()) <= set(b"ACGT$")
    r = G.random_stringset(50, 400, 32, repeat_unit=5).tobytes().split(b"$")
    assert any(len(s) > 10 and s[:5] * 2 == s[:10] or len(set(s)) == 1 for s in r)  # powers of short units
