"""Host-side file formats of the reference (suffix_array::write/read, psac -o, alphabet files, block-decomposed text):
round trips at p = 1 and across different numbers of ranks, like the reference's FileIO test (test/test_psac.cpp:306-347)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from psac_b200 import api, fileio
from psac_b200 import textgen as G


def test_write_read_roundtrip_and_raw_layout(tmp_path):
    t = G.random_dna(5003, 9)
    exp = O.construct(t, 64, 0, True)
    base = str(tmp_path / "idx")
    for dt, ib in ((np.uint32, 4), (np.uint64, 8)):
        fileio.write_suffix_array(base, exp["sa"].astype(dt), exp["lcp"].astype(dt), text=t)
        assert os.path.getsize(base + ".sa") == t.size * ib  # raw little-endian index_t, nothing else
        assert (np.fromfile(base + ".sa", dtype="<u%d" % ib) == exp["sa"]).all()
        r = fileio.read_suffix_array(base, ib, with_lcp=True)
        assert r["n"] == t.size and (r["sa"] == exp["sa"]).all() and (r["lcp"] == exp["lcp"]).all()
        lut, sigma, _ = O.alphabet(t)
        assert r["sigma"] == sigma and (r["lut"] == lut).all()
        assert open(base + ".alpha", "rb").read() == b"ACGT"


@pytest.mark.parametrize("pw,pr", [(1, 3), (4, 1), (3, 5), (13, 4)])
def test_written_by_p_ranks_read_by_q_ranks(tmp_path, pw, pr):
    n = 1000 + pw
    sa = np.random.default_rng(pw).permutation(n).astype(np.uint64)
    f = str(tmp_path / "x.sa")
    for r in range(pw):
        s, m = api.blk_dist(n, pw, r)
        fileio.write_dist_int_array(f, sa[s:s + m], r, pw, n)
    got = np.concatenate([fileio.read_dist_int_array(f, np.uint64, r, pr)[0] for r in range(pr)])
    assert (got == sa).all()


def test_psac_cli_files_and_text_blocks(tmp_path):
    t = G.random_bytes(777, 3)
    p = str(tmp_path / "text.bin")
    t.tofile(p)
    blocks = [fileio.file_block_decompose(p, r, 4) for r in range(4)]
    assert [b.size for b in blocks] == [api.blk_dist(777, 4, r)[1] for r in range(4)]
    assert (np.concatenate(blocks) == t).all()
    sa32 = np.arange(10, dtype=np.uint32)[::-1].copy()
    fileio.write_psac_cli_output(str(tmp_path / "out"), sa32, lcp=np.zeros(10, np.uint32))
    assert (np.fromfile(str(tmp_path / "out.sa64"), dtype="<u8") == sa32).all()
    assert os.path.getsize(str(tmp_path / "out.lcp64")) == 80


def test_alphabet_file_of_all_256_bytes(tmp_path):
    t = G.random_bytes_config4(70000, 1)
    f = str(tmp_path / "a.alpha")
    fileio.write_alphabet(f, text=t)
    assert os.path.getsize(f) == 256
    lut, sigma = fileio.read_alphabet(f)
    elut, esigma, _ = O.alphabet(t)
    assert sigma == esigma == 256 and (lut == elut).all() and lut[255] == 0  # the reference's 8-bit overflow


def test_size_mismatch_is_an_error(tmp_path):
    base = str(tmp_path / "bad")
    fileio.write_dist_int_array(base + ".sa", np.arange(5, dtype=np.uint64))
    fileio.write_dist_int_array(base + ".lcp", np.arange(4, dtype=np.uint64))
    with pytest.raises(api.PsacError):
        fileio.read_suffix_array(base, 8, with_lcp=True)


def test_cli_fails_loudly_without_gpu(tmp_path):
    # the psac-like command line tool (reference src/psac.cpp): flags parse, and without a GPU it reports the library's
    # error instead of computing anything on the CPU
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "psac_b200.cli", "-r", "1000", "-s", "3", "-l", "-c", "-o", str(tmp_path / "o")], cwd=root,
                       capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr
    assert not os.path.exists(str(tmp_path / "o.sa64"))


def test_cli_check_function_detects_errors():
    from psac_b200 import cli
    t = G.random_dna(3000, 4)
    exp = O.construct(t, 64, 0, True)
    assert cli._check(t, exp["sa"], exp["isa"], exp["lcp"]) is None
    bad = exp["sa"].copy()
    bad[[5, 6]] = bad[[6, 5]]
    assert cli._check(t, bad, exp["isa"], None) is not None
    lcp = exp["lcp"].copy()
    lcp[1:] += 1
    assert "LCP" in cli._check(t, exp["sa"], exp["isa"], lcp)


# ------------------------------------------------------------------------------------------- against the reference's own write / read
@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libpsacref.so not built")
@pytest.mark.parametrize("ib", [4, 8])
def test_files_written_by_the_unmodified_reference_are_read_here_and_vice_versa(tmp_path, ib):
    """suffix_array::write / read of the UNMODIFIED reference (oracle/_ref; include/suffix_array.hpp:232-265 over the MPI shim's
    stdio-backed MPI_File) against psac_b200/fileio.py: byte-identical files, and each side reads what the other wrote."""
    dt = np.uint32 if ib == 4 else np.uint64
    for t in (G.random_dna(4099, 5), G.repeats_text(300, 2), np.frombuffer(b"mississippi", np.uint8), ("ACGT" * 64).encode()):
        t = np.frombuffer(bytes(t), np.uint8) if isinstance(t, bytes) else t
        exp = O.construct(t, 64, 0, True)
        ref_base, own_base = str(tmp_path / "ref"), str(tmp_path / "own")
        O.ref_write(t, ib, ref_base)
        fileio.write_suffix_array(own_base, exp["sa"].astype(dt), exp["lcp"].astype(dt), text=t)
        for ext in (".sa", ".lcp", ".alpha"):
            assert open(ref_base + ext, "rb").read() == open(own_base + ext, "rb").read(), ext
        r = fileio.read_suffix_array(ref_base, ib, with_lcp=True)  # reference wrote, we read
        assert r["n"] == t.size and (r["sa"] == exp["sa"]).all() and (r["lcp"] == exp["lcp"]).all()
        assert (r["lut"] == O.alphabet(t)[0]).all()
        rr = O.ref_read(own_base, ib, t.size + 8)  # we wrote, the reference reads
        assert rr["n"] == t.size and (rr["sa"] == exp["sa"]).all() and (rr["lcp"] == exp["lcp"]).all()
        assert (rr["lut"] == O.alphabet(t)[0]).all() and rr["sigma"] == O.alphabet(t)[1]


def test_text_of_exactly_256_characters_writes_its_alphabet_not_a_table(tmp_path):
    f = str(tmp_path / "a.alpha")
    fileio.write_alphabet(f, text=np.frombuffer(b"ACGT" * 64, np.uint8))
    assert open(f, "rb").read() == b"ACGT"
    lut = np.zeros(256, np.uint8)
    lut[[65, 67, 71, 84]] = [1, 2, 3, 4]
    fileio.write_alphabet(f, lut=lut)
    assert open(f, "rb").read() == b"ACGT"
    with pytest.raises(api.PsacError):
        fileio.write_alphabet(f, np.zeros(3, np.uint8), np.zeros(256, np.uint8))


def test_rewriting_a_shorter_array_truncates_the_file(tmp_path):
    f = str(tmp_path / "x.sa")
    fileio.write_dist_int_array(f, np.arange(100, dtype=np.uint64))
    fileio.write_dist_int_array(f, np.arange(10, dtype=np.uint64))
    assert os.path.getsize(f) == 80 and fileio.read_dist_int_array(f, np.uint64)[1] == 10
