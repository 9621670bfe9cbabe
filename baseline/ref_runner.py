"""Runs the UNMODIFIED reference (staged under baseline/_ref by baseline/install_ref.sh) on the host CPU.

Used only by `bench.py` (the `--impl reference` arm and the `cpu_baseline` leg).  Nothing under
`dominantsparseeigenad_b200/` imports this file.

What runs is the reference's own code path, exactly as `examples/TFIM/E0.py:53-67` / `chiF.py:40-53` drive it:

    model = TFIM(N, cpu)                                   examples/TFIM/TFIM.py (verbatim copy in _ref/ref_examples)
    symeig.setDominantSparseSymeig(model.H, model.Hadjoint_to_gadjoint)
    E0, psi0 = symeig.DominantSparseSymeig.apply(model.g, k, model.dim, cpu)
    dE0, = torch.autograd.grad(E0, model.g[, create_graph=True])      ... second derivatives for chiF

with two compatibility shims and no source edit: `torch.symeig` (removed in torch >= 2.0,
Lanczos.py:98) is mapped to `torch.linalg.eigh`, and the package is loaded under a private module name so
that it cannot collide with this repository's `DominantSparseEigenAD` import alias.

If baseline/_ref is missing (it is git-ignored), `available()` is False and bench.py falls back to the
oracle port (`cpu_baseline.kind = "port"`).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
PKG_DIR = os.path.join(REF_ROOT, "DominantSparseEigenAD")
TFIM_FILE = os.path.join(REF_ROOT, "ref_examples", "TFIM.py")
_PRIVATE = "_dsea_reference_pkg"
_cache = {}


def available() -> bool:
    return os.path.exists(os.path.join(PKG_DIR, "symeig.py")) and os.path.exists(TFIM_FILE)


def _shims() -> None:
    # torch >= 2.0 keeps `torch.symeig` only as a stub that raises; Lanczos.py:98 calls it.
    torch.symeig = lambda A, eigenvectors=True: torch.linalg.eigh(A)


def load():
    """Returns (symeig module, TFIM class) of the staged reference."""
    if "symeig" in _cache:
        return _cache["symeig"], _cache["TFIM"]
    if not available():
        raise RuntimeError("baseline/_ref is not staged; run baseline/install_ref.sh")
    _shims()
    spec = importlib.util.spec_from_file_location(_PRIVATE, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[_PRIVATE] = pkg
    spec.loader.exec_module(pkg)
    symeig = importlib.import_module(_PRIVATE + ".symeig")
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)          # the example's docstrings hold "\s", "\p"
    tspec = importlib.util.spec_from_file_location(_PRIVATE + "_tfim_example", TFIM_FILE)
    tmod = importlib.util.module_from_spec(tspec)
    tspec.loader.exec_module(tmod)
    _cache["symeig"], _cache["TFIM"] = symeig, tmod.TFIM
    return symeig, tmod.TFIM


def required_host_bytes(N: int, k: int) -> int:
    """Peak host memory of one reference solve: the (n, k) basis (Lanczos.py:49), the all-Ritz-vector
    product of the same size (Lanczos.py:99), the (n, N) int64 flip table (TFIM.py:48-51), one (n, N) fp64 gather
    temporary per operator call (TFIM.py:97) and the ~5 (n, N) int64 temporaries of `_diags` (TFIM.py:39-46)."""
    n = 1 << N
    return 2 * 8 * n * k + 8 * n * N * 2 + max(5 * 8 * n * N, 8 * n * 16)


class _TimedCalls:
    def __init__(self, fn):
        self.fn, self.calls, self.seconds = fn, 0, 0.0

    def __call__(self, v):
        t = time.perf_counter()
        out = self.fn(v)
        self.seconds += time.perf_counter() - t
        self.calls += 1
        return out


_models = {}


def model_for(N: int):
    if N not in _models:
        _, TFIM = load()
        with contextlib.redirect_stdout(io.StringIO()):          # the constructor prints two lines
            _models.clear()                                      # one flip table at a time
            _models[N] = TFIM(N, torch.device("cpu"))
    return _models[N]


def solve(N: int, k: int, g: float, mode: str = "E0_dE0", seed: int = 1234) -> dict:
    """One forward+backward solve with the reference; wall-clock split like SURVEY 8d asks.

    mode "E0_dE0": E0 and dE0/dg (E0.py:60-63 without the second derivative);
         "E0_d2E0": E0, dE0/dg, d2E0/dg2 (E0.py:60-64);
         "chiF":  fidelity susceptibility (chiF.py:46-53)."""
    symeig, _ = load()
    model = model_for(N)
    model.g = torch.tensor([g], dtype=torch.float64, requires_grad=True)
    H = _TimedCalls(model.H)
    torch.manual_seed(seed)
    symeig.setDominantSparseSymeig(H, model.Hadjoint_to_gadjoint)
    out = {"N": N, "k": k, "g": g, "mode": mode}
    t0 = time.perf_counter()
    E0, psi0 = symeig.DominantSparseSymeig.apply(model.g, k, model.dim, torch.device("cpu"))
    t1 = time.perf_counter()
    out.update(fwd=t1 - t0, calls_fwd=H.calls, mv_fwd=H.seconds, E0=E0.item())
    if mode == "E0_dE0":
        dE0, = torch.autograd.grad(E0, model.g)
        out["dE0"] = dE0.item()
    elif mode == "E0_d2E0":
        dE0, = torch.autograd.grad(E0, model.g, create_graph=True)
        d2E0, = torch.autograd.grad(dE0, model.g)
        out["dE0"], out["d2E0"] = dE0.item(), d2E0.item()
    elif mode == "chiF":
        logF = torch.log(psi0.detach().matmul(psi0))
        dlogF, = torch.autograd.grad(logF, model.g, create_graph=True)
        d2logF, = torch.autograd.grad(dlogF, model.g)
        out["chiF"] = -d2logF.item()
    else:
        raise ValueError(mode)
    t2 = time.perf_counter()
    out.update(bwd=t2 - t1, total=t2 - t0, calls_bwd=H.calls - out["calls_fwd"], mv_bwd=H.seconds - out["mv_fwd"])
    del E0, psi0
    return out


def warm(N: int) -> None:
    """Builds the model tables and applies H once (the first application is 4-5x slower: page faults)."""
    model = model_for(N)
    model.g = torch.tensor([1.0], dtype=torch.float64)
    model.H(torch.randn(model.dim, dtype=torch.float64))


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--spins", type=int, default=20)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--g", type=float, default=1.0)
    ap.add_argument("--mode", default="E0_dE0")
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    torch.set_num_threads(a.threads or (os.cpu_count() or 1))
    warm(a.spins)
    r = solve(a.spins, a.k, a.g, a.mode)
    r["threads"] = torch.get_num_threads()
    print(json.dumps(r))
