"""Sharded construction on >= 2 GPUs of one box: parity of the block-distributed SA / ISA / LCP with the CPU oracle.
Runs tests/sharded_worker.py under torchrun, one rank per GPU (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_construct_matches_oracle(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port",
           str(29500 + world), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]


@pytest.mark.parametrize("gpus", [2, 4])
def test_one_process_several_gpus_behind_the_reference_class(gpus, tmp_path):
    """psacb200_multi_*: whole host arrays in and out, the GPUs sharded inside the engine (threads + plain peer access):
    the Python binding against the oracle, and the C++ mirror of the reference class against its own one-GPU result."""
    if _gpus() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    import numpy as np
    from oracle import pyoracle as O
    from psac_b200 import api, textgen as G
    m = api.MultiEngine(gpus)
    for t, ib, lcp, k in [(G.random_dna((gpus << 19) + 3, 71), 8, True, 0), (G.random_dna(gpus << 18, 72), 4, False, 6),
                          (G.repeats_text(30000 * gpus, 73), 8, True, 0), (G.random_dna(5000, 74), 8, True, 0)]:
        r = m.construct(t, ib, lcp, k)
        exp = O.construct(t, 64, k, lcp)
        assert (r["sa"] == exp["sa"]).all() and (r["isa"] == exp["isa"]).all()
        if lcp:
            assert (r["lcp"] == exp["lcp"]).all()
    m.close()
    exe = str(tmp_path / "test_shim_multi")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim_multi.cpp"),
                           "-o", exe, "-L", os.path.join(ROOT, "psac_b200"), "-lpsacb200", "-Wl,-rpath," + os.path.join(ROOT, "psac_b200")])
    r = subprocess.run([exe, str(gpus)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
