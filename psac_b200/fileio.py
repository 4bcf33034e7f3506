"""On-disk formats of the reference: host-side only, no GPU involved.

* ``<base>.sa`` / ``<base>.lcp``  -- ``suffix_array::write`` / ``read`` (reference include/suffix_array.hpp:129-166, 232-265):
  the raw little-endian ``index_t`` array, ranks writing their blocks in rank order (``write_dist_int_array`` ->
  ``mxx::coll_file::write_ordered``); ``<base>.alpha`` = the characters that occur, one byte each, in code order
  (include/alphabet.hpp:296-347).
* ``<o>.sa64`` / ``<o>.lcp64``    -- the ``psac`` command line tool (reference src/psac.cpp:127-128, 142): the same raw
  arrays with ``index_t = uint64_t``.
* block-decomposed text input    -- ``mxx::file_block_decompose`` (reference ext/mxx/include/mxx/file.hpp:218-251): rank r
  reads bytes [start_r, start_r + size_r) of the file under mxx::blk_dist.
"""
import os

import numpy as np

from . import api


def write_dist_int_array(filename, local_block, rank=0, world=1, n=None):
    """Every rank writes its block at its blk_dist offset (rank 0 creates / truncates the file first; callers
    synchronise between the two steps when world > 1, see ``write_ordered_all``)."""
    a = np.ascontiguousarray(local_block)
    if a.dtype not in (np.uint32, np.uint64, np.uint8):
        raise api.PsacError("write_dist_int_array: unsigned 8/32/64-bit arrays only")
    total = a.size if n is None else int(n)
    start, size = api.blk_dist(total, world, rank)
    if size != a.size:
        raise api.PsacError("write_dist_int_array: the block is not this rank's blk_dist block")
    mode = "r+b" if os.path.exists(filename) else "w+b"
    with open(filename, mode) as f:
        if rank == 0:
            f.truncate(total * a.dtype.itemsize)  # a stale longer file would change the n that read() infers from the size
        f.seek(start * a.dtype.itemsize)
        f.write(a.tobytes())


def read_dist_int_array(filename, dtype, rank=0, world=1):
    """reference read_dist_int_array (suffix_array.hpp:139-166): n = file size / sizeof(T), rank r gets its block"""
    dt = np.dtype(dtype)
    n = os.path.getsize(filename) // dt.itemsize
    start, size = api.blk_dist(n, world, rank)
    return np.fromfile(filename, dtype=dt, count=size, offset=start * dt.itemsize), n


def file_block_decompose(filename, rank=0, world=1):
    """reference mxx::file_block_decompose: this rank's block of the text file as uint8"""
    n = os.path.getsize(filename)
    start, size = api.blk_dist(n, world, rank)
    return np.fromfile(filename, dtype=np.uint8, count=size, offset=start)


def write_alphabet(filename, text=None, lut=None):
    """``alphabet::write``: the used characters in increasing code order (= increasing byte order, alphabet.hpp:157-164).
    Pass EITHER the text (its distinct characters are written) OR a 256-entry code table as produced by psacb200_alphabet."""
    if (text is None) == (lut is None):
        raise api.PsacError("write_alphabet: pass exactly one of text= / lut=")
    if text is not None:
        chars = np.unique(np.asarray(text).astype(np.uint8))
    else:
        table = np.asarray(lut).astype(np.uint8)
        if table.size != 256:
            raise api.PsacError("write_alphabet: a code table has 256 entries")
        used = np.nonzero(table)[0]
        if used.size == 255 and table[255] == 0:  # the 8-bit table wrapped: every byte value occurs (sigma = 256)
            used = np.arange(256)
        chars = used.astype(np.uint8)
    with open(filename, "wb") as f:
        f.write(chars.tobytes())


def read_alphabet(filename):
    """``alphabet::read`` -> (lut[256], sigma): codes 1..sigma in byte order, stored in 8 bits like the reference"""
    chars = np.fromfile(filename, dtype=np.uint8)
    lut = np.zeros(256, np.uint8)
    code = 1
    for c in np.unique(chars):
        lut[int(c)] = code & 0xFF
        code += 1
    return lut, int(np.unique(chars).size)


def write_suffix_array(basename, sa, lcp=None, text=None, rank=0, world=1, n=None):
    """``suffix_array::write(basename)`` (suffix_array.hpp:232-243): .sa, .lcp (if built), .alpha"""
    write_dist_int_array(basename + ".sa", sa, rank, world, n)
    if lcp is not None:
        write_dist_int_array(basename + ".lcp", lcp, rank, world, n)
    if text is not None and rank == 0:
        write_alphabet(basename + ".alpha", text=text)


def read_suffix_array(basename, index_bytes=8, with_lcp=False, rank=0, world=1):
    """``suffix_array::read(basename)`` (suffix_array.hpp:245-265): returns dict(sa, lcp, lut, sigma, n)"""
    dt = np.uint32 if index_bytes == 4 else np.uint64
    sa, n = read_dist_int_array(basename + ".sa", dt, rank, world)
    lcp = None
    if with_lcp:
        lcp, n2 = read_dist_int_array(basename + ".lcp", dt, rank, world)
        if n2 != n:
            raise api.PsacError("SA and LCP have to have same size")  # the reference's message (:250-252)
    lut, sigma = read_alphabet(basename + ".alpha") if os.path.exists(basename + ".alpha") else (None, 0)
    return dict(sa=sa, lcp=lcp, lut=lut, sigma=sigma, n=n)


def write_psac_cli_output(prefix, sa, lcp=None, rank=0, world=1, n=None):
    """what ``psac -o <prefix>`` writes: <prefix>.sa64 and, with -l, <prefix>.lcp64 (src/psac.cpp:127-128, 142)"""
    write_dist_int_array(prefix + ".sa64", np.ascontiguousarray(sa, np.uint64), rank, world, n)
    if lcp is not None:
        write_dist_int_array(prefix + ".lcp64", np.ascontiguousarray(lcp, np.uint64), rank, world, n)
