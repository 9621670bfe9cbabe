"""Lanczos eigensolver — same public names as the reference's DominantSparseEigenAD/Lanczos.py.

    Lanczos(A, k, device, *, sparse=False, dim=None)             -> (Qk, T)          (Lanczos.py:3-77)
    symeigLanczos(A, k, device, extreme="both", *, sparse, dim)  -> eigenpair(s)     (Lanczos.py:79-105)

`A` may be a dense symmetric torch.Tensor, a Python callable (with sparse=True, dim=n) or one of the
native operator callables (`TFIM.H`, `SparseMatrixOperator.H`).  All arithmetic runs in libdsea on the
GPU: the k-step loop with fused two-pass full re-orthogonalisation, the single-CTA tridiagonal
eigensolver and the Ritz GEMV.  Results are returned on the device of the inputs (`device` argument for
callables), so CPU callers get CPU tensors back.
"""
from __future__ import annotations

import torch

from . import _lib
from .operators import as_operator

_WHICH = {"min": _lib.DSEA_MIN, "max": _lib.DSEA_MAX, "both": _lib.DSEA_BOTH}


def _resolve(A, device, sparse, dim):
    if isinstance(A, torch.Tensor):
        return as_operator(A, None, A.device), A.device
    if not sparse and not hasattr(A, "_dsea_operator"):
        raise TypeError("A must be a torch.Tensor unless sparse=True (Lanczos.py:42-48)")
    return as_operator(A, dim, device), torch.device(device)


def Lanczos(A, k, device=torch.device("cpu"), *, sparse=False, dim=None):
    """Returns (Qk, T): Qk (n, k) with orthonormal columns, T (k, k) tridiagonal (Lanczos.py:12-15).

    Qk is a column-contiguous (n, k) view (stride (1, ldq)) of the device basis — the transpose of the
    reference's row-major storage — so `Qk.T @ A @ Qk == T` holds as documented.
    """
    op, out_dev = _resolve(A, device, sparse, dim)
    _, _, _, st = op.lanczos(None, int(k), _lib.DSEA_MIN, want_info=False)
    if st.get("basis") == "fp32":
        raise RuntimeError("Lanczos() returns the fp64 basis; switch runtime.set_basis_precision('fp64') for it")
    n, ldq = op.n_loc, st["ldq"]
    Qk = st["Q"].view(int(k), ldq)[:, :n].t()
    a, b = st["alpha"], st["beta"][: int(k) - 1]
    T = torch.diag(a) + torch.diag(b, 1) + torch.diag(b, -1)
    return Qk.to(out_dev), T.to(out_dev)


def symeigLanczos(A, k, device=torch.device("cpu"), extreme="both", *, sparse=False, dim=None):
    """Extreme eigenpair(s) of a real symmetric operator (Lanczos.py:79-105).

    extreme="both" -> (eigval_min, eigvector_min, eigval_max, eigvector_max); "min"/"max" -> one pair.
    Eigenvalues are 0-dim tensors, eigenvectors have unit norm and arbitrary sign.
    """
    if extreme not in _WHICH:
        raise ValueError("extreme must be 'min', 'max' or 'both'")
    op, out_dev = _resolve(A, device, sparse, dim)
    evals, vmin, vmax, _ = op.lanczos(None, int(k), _WHICH[extreme])
    if extreme == "min":
        return evals[0].to(out_dev), vmin.to(out_dev)
    if extreme == "max":
        return evals[1].to(out_dev), vmax.to(out_dev)
    return evals[0].to(out_dev), vmin.to(out_dev), evals[1].to(out_dev), vmax.to(out_dev)
