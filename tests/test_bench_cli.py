"""CPU: bench.py's reference arm (the unmodified reference staged under baseline/_ref — or the oracle port when
it is not staged — timed on the host cores) prints one valid JSON line whose value is a MEASUREMENT of the
workload it names, and the product arm refuses to run without a GPU instead of falling back to the CPU."""
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
    out = _run("--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-sample-spins", "10", "--cpu-sample-k", "50",
               "--ref-budget-s", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "tfim_fwd_bwd_seconds_per_solve" and d["unit"] == "s/solve"
    assert d["higher_is_better"] is False and d["dtype"] == "f64" and d["gpu_launches"] == 0
    # the product workload (N=24, k=200) was not run inside a zero budget: the line must name what WAS measured
    assert d["config"]["workload"] == "tfim_N10_k50_E0_plus_dE0dg"
    assert d["config"]["product_workload"] == "tfim_N24_k200_E0_plus_dE0dg"
    cb = d["cpu_baseline"]
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "DominantSparseEigenAD", "symeig.py"))
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1
    assert cb["value"] == d["value"] and cb["sample_workload"] == d["config"]["workload"]
    assert cb["attempt"]["attempted"] is False and "why_not" in cb["attempt"]
    # value is the measured wall time of that one solve, and ms_per_step / steps agree with it
    assert d["steps"] == 1 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-9
    assert abs(cb["detail"]["total"] - d["value"]) < 1e-12 and d["wall_seconds"] >= d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s/solve", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_the_product_workload_when_it_fits():
    """With the product workload small enough for this host the arm measures THAT and labels it so."""
    out = _run("--impl", "reference", "--spins", "12", "--k", "40", "--cpu-sample-spins", "10", "--cpu-sample-k", "30",
               "--ref-budget-s", "300")
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "DominantSparseEigenAD", "symeig.py"))
    if not staged:
        pytest.skip("baseline/_ref is not staged")
    assert d["config"]["workload"] == "tfim_N12_k40_E0_plus_dE0dg" and "product_workload" not in d["config"]
    assert d["cpu_baseline"]["attempt"]["attempted"] is True
    assert d["cpu_baseline"]["detail"]["N"] == 12 and d["cpu_baseline"]["detail"]["k"] == 40
    assert abs(d["cpu_baseline"]["detail"]["E0"] / 12 + 1.28) < 0.05        # E0/N of the g=1 chain


def test_reference_arm_other_ranks_exit_quietly():
    out = _run("--impl", "reference", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr or "CUDA" in out.stderr
