"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/psacb200.h
declares, and fails LOUDLY (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from psac_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "psacb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psacb200_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = _declared_functions()
    for must in ("psacb200_create", "psacb200_destroy", "psacb200_construct", "psacb200_construct_device", "psacb200_construct_alphabet",
                 "psacb200_alphabet", "psacb200_sort_pairs", "psacb200_last_error", "psacb200_get_stats", "psacb200_launch_count"):
        assert must in names


def test_library_exports_every_declared_symbol():
    L = api.lib()
    for name in _declared_functions():
        assert hasattr(L, name), "libpsacb200.so does not export " + name


def test_stats_struct_matches_header_size():
    # 4 u64-ish header words + floats; the C side memsets sizeof(psacb200_stats) -- keep the mirror in sync
    assert C.sizeof(api.Stats) == 120  # static_assert'ed on the C side (engine.cu)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only boxes")
    with pytest.raises(api.PsacError) as ei:
        api.Engine(0)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(api.PsacError):
        api.SuffixArray(4).construct(b"mississippi")


def test_null_engine_is_an_error_not_a_crash():
    L = api.lib()
    sa = np.zeros(4, np.uint32)
    rc = L.psacb200_construct(None, None, 0, 4, 0, 0, sa.ctypes.data_as(C.c_void_p), None, None)
    assert rc < 0 and b"null" in L.psacb200_last_error()
