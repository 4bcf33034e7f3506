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
    assert C.sizeof(api.Stats) == 128  # static_assert'ed on the C side (engine.cu)


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


def test_new_entry_points_reject_a_null_engine():
    L = api.lib()
    buf = np.zeros(8, np.uint64)
    n_out = C.c_uint64(7)
    rc = L.psacb200_construct_ss(None, None, 0, C.c_uint8(36), 8, 0, None, buf.ctypes.data_as(C.c_void_p), None, None, C.byref(n_out))
    assert rc < 0 and b"null" in L.psacb200_last_error()
    nd = C.c_uint32(0)
    rc = L.psacb200_construct_wide(None, None, 0, 4, 1, 8, 0, 0, buf.ctypes.data_as(C.c_void_p), None, None, None, C.byref(nd))
    assert rc < 0 and b"null" in L.psacb200_last_error()


def _build_cpp_shim_test(tmp_path):
    import subprocess
    exe = str(tmp_path / "test_shim")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                           "-L", os.path.join(ROOT, "psac_b200"), "-lpsacb200", "-Wl,-rpath," + os.path.join(ROOT, "psac_b200"), "-o", exe])
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    # include/psacb200/suffix_array.hpp: the C++ mirror of the reference class (reference tests: test/test_psac.cpp:101-176)
    import subprocess
    import torch
    api.lib()
    exe = _build_cpp_shim_test(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the shim's results are checked by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3, r.stderr  # std::runtime_error("... no CUDA device ...") -- no CPU fallback


@pytest.mark.gpu
def test_cpp_shim_matches_reference_goldens_on_gpu(tmp_path):
    import subprocess
    exe = _build_cpp_shim_test(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_header_is_valid_c99_and_links(tmp_path):
    import subprocess
    api.lib()
    exe = str(tmp_path / "test_header")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_header.c"),
                           "-L", os.path.join(ROOT, "psac_b200"), "-lpsacb200", "-Wl,-rpath," + os.path.join(ROOT, "psac_b200"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_shim_file_formats(tmp_path):
    import subprocess
    api.lib()
    exe = str(tmp_path / "test_shim_io")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim_io.cpp"),
                           "-L", os.path.join(ROOT, "psac_b200"), "-lpsacb200", "-Wl,-rpath," + os.path.join(ROOT, "psac_b200"), "-o", exe])
    r = subprocess.run([exe, str(tmp_path / "idx")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # the files are the reference's formats: the Python side reads them back
    from psac_b200 import fileio
    got = fileio.read_suffix_array(str(tmp_path / "idx"), 4, with_lcp=True)
    assert got["sa"].tolist() == [10, 7, 4, 1, 0, 9, 8, 6, 3, 5, 2] and got["sigma"] == 4
