"""Deterministic synthetic texts for parity tests and the benchmark.

A counter-based generator (splitmix64 of seed + block index), so any shard can generate
its own block and the same bytes are reproducible on any machine -- the reference's own
``rand_dna`` depends on glibc ``rand()`` (include/alphabet.hpp:32-45).
"""
import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed, idx):
    """splitmix64 output for counters ``idx`` (uint64 array) and stream ``seed``."""
    with np.errstate(over="ignore"):
        z = (idx.astype(np.uint64) + np.uint64(1)) * _GOLDEN + np.uint64(seed) * _M1
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def random_bytes(n, seed, start=0, chunk=1 << 24):
    """n uniform bytes: byte i = byte (i % 8) of splitmix64(seed, i // 8)."""
    out = np.empty(n, np.uint8)
    pos = 0
    while pos < n:
        m = min(chunk, n - pos)
        g0 = start + pos
        w0 = g0 // 8
        w1 = (g0 + m + 7) // 8
        words = splitmix64(seed, np.arange(w0, w1, dtype=np.uint64))
        b = words.view(np.uint8)  # little-endian host
        off = g0 - w0 * 8
        out[pos:pos + m] = b[off:off + m]
        pos += m
    return out


def random_dna(n, seed, start=0, chunk=1 << 24):
    """n uniform characters over ACGT: char i = "ACGT"[(splitmix64(seed, i // 32) >> 2*(i % 32)) & 3]."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    out = np.empty(n, np.uint8)
    pos = 0
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    while pos < n:
        m = min(chunk, n - pos)
        g0 = start + pos
        w0 = g0 // 32
        w1 = (g0 + m + 31) // 32
        words = splitmix64(seed, np.arange(w0, w1, dtype=np.uint64))
        codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
        off = g0 - w0 * 32
        out[pos:pos + m] = acgt[codes[off:off + m]]
        pos += m
    return out


def random_bytes_config4(n, seed):
    """BASELINE config 4 text: uniform bytes with the last byte forced != 0xFF."""
    t = random_bytes(n, seed)
    if n and t[-1] == 0xFF:
        t[-1] = 0xFE
    return t


def repeats_text(nwords, seed):
    """Highly repetitive text in the spirit of test/test_psac.cpp:178-190 (own PRNG instead of std::rand)."""
    words = [b"helloworld", b"blahlablah", b"ellow", b"worldblah", b"rld", b"hello"]
    pick = splitmix64(seed, np.arange(nwords, dtype=np.uint64)) % np.uint64(len(words))
    return np.frombuffer(b"".join(words[int(i)] for i in pick), np.uint8).copy()


def periodic_text(unit, reps):
    """(unit)^reps, e.g. (abc)^n of test/test_suffixtree.cpp:131-162."""
    return np.frombuffer(bytes(unit) * reps, np.uint8).copy()


# ---------------------------------------------------------------- same generators as torch ops (any device)
def random_stringset(nstrings, max_len, seed, alphabet=b"ACGT", sep=b"$", repeat_unit=0):
    """Flat text of `nstrings` strings (lengths 1..max_len, own PRNG) separated by `sep` -- the input of the reference's
    simple_dstringset (include/stringset.hpp:33-152).  repeat_unit > 0: every string is a power of a short random unit
    (many identical suffixes across strings, long common prefixes)."""
    al = np.frombuffer(alphabet, np.uint8)
    r = splitmix64(seed, np.arange(nstrings, dtype=np.uint64))
    lens = (r % np.uint64(max_len)).astype(np.int64) + 1
    out = []
    for i in range(nstrings):
        L = int(lens[i])
        if repeat_unit > 0:
            u = int(splitmix64(seed + 7, np.array([i], np.uint64))[0] % np.uint64(repeat_unit)) + 1
            unit = al[(splitmix64(seed + 11 + i, np.arange(u, dtype=np.uint64)) % np.uint64(al.size)).astype(np.int64)]
            s = np.tile(unit, L // u + 1)[:L]
        else:
            s = al[(splitmix64(seed + 13 + i, np.arange(L, dtype=np.uint64)) % np.uint64(al.size)).astype(np.int64)]
        out.append(s.tobytes())
    return np.frombuffer(sep.join(out), np.uint8).copy()


def _t_lsr(z, k):
    """logical right shift of int64 tensors (torch's >> is arithmetic)"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _t_wrap(x):
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def splitmix64_torch(seed, idx):
    """splitmix64 of psac_b200.textgen.splitmix64 on int64 tensors (two's complement wrap-around arithmetic)."""
    z = (idx + 1) * _t_wrap(int(_GOLDEN)) + _t_wrap(int(seed) * int(_M1))
    z = (z ^ _t_lsr(z, 30)) * _t_wrap(int(_M1))
    z = (z ^ _t_lsr(z, 27)) * _t_wrap(int(_M2))
    return z ^ _t_lsr(z, 31)


def random_dna_torch(n, seed, device, start=0, chunk=1 << 26):
    """Bit-identical to random_dna(n, seed, start) but generated with torch ops on `device`; returns a uint8 tensor."""
    import torch
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    shifts = (torch.arange(32, dtype=torch.int64, device=device) * 2)[None, :]
    pos = 0
    while pos < n:
        m = min(chunk, n - pos)
        g0 = start + pos
        w0, w1 = g0 // 32, (g0 + m + 31) // 32
        words = splitmix64_torch(seed, torch.arange(w0, w1, dtype=torch.int64, device=device))
        codes = ((words[:, None] >> shifts) & 3).reshape(-1)
        off = g0 - w0 * 32
        out[pos:pos + m] = acgt[codes[off:off + m]]
        pos += m
    return out


def random_bytes_torch(n, seed, device, start=0, chunk=1 << 29):
    """Bit-identical to random_bytes(n, seed, start), generated with torch ops on `device`; returns a uint8 tensor."""
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    pos = 0
    while pos < n:
        m = min(chunk, n - pos)
        g0 = start + pos
        w0, w1 = g0 // 8, (g0 + m + 7) // 8
        words = splitmix64_torch(seed, torch.arange(w0, w1, dtype=torch.int64, device=device))
        b = words.view(torch.uint8)  # little-endian bytes of every word, as in the numpy generator
        off = g0 - w0 * 8
        out[pos:pos + m] = b[off:off + m]
        pos += m
    return out
