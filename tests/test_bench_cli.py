"""CPU: bench.py's reference arm (the oracle timed on the host cores) prints one valid JSON line, and the
product arm refuses to run without a GPU instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env={**os.environ, **(env or {})})


def test_reference_arm_prints_contract_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-spins", "10")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "tfim_fwd_bwd_seconds_per_solve" and d["unit"] == "s/solve"
    assert d["higher_is_better"] is False and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "tfim_N24_k200_E0_plus_dE0dg"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "N=10" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "s/solve", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = _run("--impl", "reference", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr or "CUDA" in out.stderr
