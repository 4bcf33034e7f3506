"""bench.py contract on a CPU-only box: the reference arm (`--impl reference`) times the unmodified reference (oracle/_ref,
or the plain-C port where that is absent) on host cores and prints ONE JSON line with the keys the driver reads; ranks other
than 0 do no work; the B200 arm fails loudly without a GPU (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.parametrize("config,metric", [(2, "suffixes/sec SA+LCP build"), (1, "suffixes/sec SA build"), (5, "suffixes/sec SA+LCP+suffix-tree build"),
                                           (6, "suffixes/sec generalized SA+LCP build (string set)")])
def test_reference_arm_prints_the_contract_line(config, metric):
    r = _run(["--impl", "reference", "--config", str(config), "--steps", "1", "--warmup", "0", "--cpu-log2n", "16", "--gpus", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "suffixes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_no_work():
    r = _run(["--impl", "reference", "--steps", "1", "--cpu-log2n", "14", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for cfg in ("2", "6"):
        r = _run(["--config", cfg, "--steps", "1", "--no-cpu-baseline"])
        assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
