"""Dominant non-symmetric eigensolver primitives — same names as the reference's eig.py.

    DominantEig.apply(A, k, which="LM")                      -> (eigval[1], lefteigvector, righteigvector)   (eig.py:5-62)
    setDominantSparseEig(A, AT, Aadjoint_to_gadjoint)
    DominantSparseEig.apply(g, k)                            -> same triple                                   (eig.py:64-152)

The reference delegates all arithmetic to scipy on the CPU: two ARPACK `eigs(k=1, ncv=k)` calls (A and
A^T) forward, two `gmres(tol=1e-12)` solves backward.  Here both are Krylov loops on the GPU built from
the same libdsea kernels as the symmetric path:

    forward   explicitly restarted k-vector Arnoldi for A and for A^T (operator application: dense GEMV
              kernel, or the user's callable); Gram-Schmidt = fused two-pass GEMV, two sweeps; the k x k
              Hessenberg Ritz problem (tiny) goes to LAPACK through torch.linalg.eig; Ritz vector = GEMV;
    backward  restarted GMRES on (A - lambda I) and (A^T - lambda I) with the same Arnoldi kernels, the
              (m+1) x m least-squares problem solved by torch.linalg.lstsq; adjoint = three outer products
              (eig.py:58-60, 145-147).

Conventions follow eig.py:22-24: l^T r = 1 and r^T r = 1; the eigenvalue must be real.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from . import _lib, runtime
from .operators import DenseOperator
from .runtime import F64, context, dev_vec, empty, ptr, stream_ptr


# -------------------------------------------------------------------------------------------------
# level-1 pieces through libdsea (deterministic two-stage reductions; results stay on the device)
# -------------------------------------------------------------------------------------------------
def _dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a . b as a 1-element device tensor (dsea_dot); no host synchronisation."""
    rt = context()
    out = torch.empty(1, dtype=F64, device=rt.device)
    _lib.check(rt.lib.dsea_dot(rt.handle, a.numel(), ptr(a), ptr(b), out.data_ptr(), stream_ptr()))
    return out


def _axpby(a, x: torch.Tensor, b, y: torch.Tensor) -> torch.Tensor:
    """y <- a x + b y in place (dsea_axpby); a, b are 1-element device tensors, python floats or None (= 1)."""
    rt = context()
    def dev_scalar(s):
        if s is None or isinstance(s, torch.Tensor):
            return s
        return torch.tensor([float(s)], dtype=F64, device=rt.device)
    a_, b_ = dev_scalar(a), dev_scalar(b)
    _lib.check(rt.lib.dsea_axpby(rt.handle, y.numel(), None if a_ is None else a_.data_ptr(), ptr(x),
                                 None if b_ is None else b_.data_ptr(), ptr(y), stream_ptr()))
    return y


def _normalised(x: torch.Tensor) -> torch.Tensor:
    return _axpby(0.0, x, torch.rsqrt(_dot(x, x)), x)          # x <- x / |x| (one reduction + one pass, no sync)

GMRES_RTOL = 1e-12      # eig.py:54,57
GMRES_ATOL = 1e-12
EIG_RTOL = 1e-13        # ARPACK is called with tol=0 (machine precision)
MAX_RESTARTS = 200


# -------------------------------------------------------------------------------------------------
# Krylov machinery on top of the C ABI
# -------------------------------------------------------------------------------------------------
def _arnoldi(apply: Callable[[torch.Tensor], torch.Tensor], n: int, m: int, v0: torch.Tensor):
    """m-step Arnoldi from v0.  Returns (Q buffer, ldq, Hbar (m+1, m) on the host, |v0|)."""
    rt = context()
    lib = rt.lib
    ldq = rt.col_stride(n)
    Q = empty((m + 1) * ldq, rt.device)
    Q[:n].copy_(v0)
    H = torch.zeros((m, m + 1), dtype=F64, device=rt.device)           # row j of this tensor = column j of Hbar
    norm2 = torch.empty(1, dtype=F64, device=rt.device)
    st = stream_ptr()
    _lib.check(lib.dsea_arnoldi_start(rt.handle, n, ptr(Q), norm2.data_ptr(), st))
    for i in range(m):
        u = apply(Q[i * ldq:i * ldq + n])
        _lib.check(lib.dsea_arnoldi_step(rt.handle, n, m, i, ptr(Q), ptr(u), H.data_ptr(), st))
    Hbar = H.t().cpu()                                                   # one small D2H copy, synchronises
    return Q, ldq, Hbar, float(norm2.sqrt().item())


def _combine(Q: torch.Tensor, n: int, m: int, coef: torch.Tensor, add: torch.Tensor | None = None) -> torch.Tensor:
    """add + sum_j coef[j] Q[:, j] through the Ritz GEMV kernel."""
    rt = context()
    out = empty(n, rt.device)
    c = coef.to(device=rt.device, dtype=F64).contiguous()
    _lib.check(rt.lib.dsea_combine(rt.handle, n, m, ptr(Q), ptr(c), ptr(add), ptr(out), stream_ptr()))
    return out


def _select(w: torch.Tensor, which: str) -> int:
    if which == "LM":
        return int(torch.argmax(w.abs()))
    if which == "SM":
        return int(torch.argmin(w.abs()))
    if which == "LR":
        return int(torch.argmax(w.real))
    if which == "SR":
        return int(torch.argmin(w.real))
    raise ValueError("which must be one of 'LM', 'SM', 'LR', 'SR' (eig.py:16-19)")


def dominant_eigpair(apply, n: int, k: int, which: str = "LM"):
    """Explicitly restarted Arnoldi(k) for one eigenpair; returns (eigval float, unit eigenvector)."""
    rt = context()
    k = max(1, min(int(k), n))
    v = runtime.start_vector(n, "lanczos")
    lam, x, converged, resid = None, None, False, float("inf")
    for _ in range(MAX_RESTARTS):
        Q, ldq, Hbar, _ = _arnoldi(apply, n, k, v)
        # an (almost) invariant subspace shows up as a tiny sub-diagonal entry: truncate there
        sub = torch.diagonal(Hbar, offset=-1)[:k]
        scale_h = Hbar.abs().max().item()
        small = (sub.abs() <= 1e-12 * scale_h).nonzero()
        m = int(small[0]) + 1 if small.numel() else k
        w, Y = torch.linalg.eig(Hbar[:m, :m])
        j = _select(w, which)
        is_complex = abs(w[j].imag.item()) > 1e-8 * max(abs(w[j].item()), 1e-300)
        lam = w[j].real.item()
        # a not-yet-converged complex-conjugate Ritz pair may transiently lead: restart from the real part of
        # its Ritz vector and only insist on a real eigenvalue once the residual has converged (eig.py:31
        # asserts on ARPACK's converged result only)
        y = Y[:, j].real.clone()
        y /= y.norm()
        x = _combine(Q, n, m, y)
        resid = abs(Hbar[m, m - 1].item() * Y[m - 1, j].abs().item()) if m < Hbar.shape[0] else 0.0
        converged = m < k or resid <= EIG_RTOL * max(abs(w[j].item()), 1e-300)
        if converged:
            if is_complex:
                raise AssertionError("The desired eigenvalue of the matrix must be real")      # eig.py:31
            break
        v = x
    if not converged:
        import warnings
        warnings.warn(f"restarted Arnoldi(k={k}) stopped after {MAX_RESTARTS} restarts with residual {resid:.3e}",
                      _lib.ConvergenceWarning, stacklevel=2)
    x = _normalised(x)
    # reproducible sign: largest-magnitude component positive
    if x[torch.argmax(x.abs())] < 0:
        x = -x
    return lam, x


def gmres_solve(apply, b: torch.Tensor, restart: int = 64, rtol: float = GMRES_RTOL, atol: float = GMRES_ATOL,
                maxiter: int = 200) -> torch.Tensor:
    """Restarted GMRES from x0 = 0 (scipy's default), stopping at |r| <= max(rtol |b|, atol)."""
    rt = context()
    n = b.numel()
    b = dev_vec(b, rt.device)
    bnorm = float(torch.sqrt(_dot(b, b)).item())
    x = torch.zeros(n, dtype=F64, device=rt.device)
    target = max(rtol * bnorm, atol)
    if bnorm <= target:
        return x
    r = b.clone()
    m = max(1, min(restart, n))
    for _ in range(maxiter):
        Q, ldq, Hbar, beta = _arnoldi(apply, n, m, r)
        sub = torch.diagonal(Hbar, offset=-1)[:m]
        small = (sub.abs() <= 1e-12 * Hbar.abs().max().item()).nonzero()
        mm = int(small[0]) + 1 if small.numel() else m
        rhs = torch.zeros(mm + 1, 1, dtype=F64)
        rhs[0, 0] = beta
        y = torch.linalg.lstsq(Hbar[:mm + 1, :mm], rhs, driver="gelsd").solution[:, 0]
        x = _combine(Q, n, mm, y, x)
        r = _axpby(-1.0, apply(x), None, b.clone())                        # r = b - A x
        rnorm = float(torch.sqrt(_dot(r, r)).item())                       # the one host sync of the cycle
        if rnorm <= target:
            break
    else:
        import warnings
        warnings.warn(f"GMRES({m}) stopped after {maxiter} cycles with |r| = {rnorm:.3e} > {target:.3e}",
                      _lib.ConvergenceWarning, stacklevel=2)
    return x


# -------------------------------------------------------------------------------------------------
# operator plumbing
# -------------------------------------------------------------------------------------------------
def _dense_apply(op: DenseOperator, shift: float = 0.0):
    sh = None if shift == 0.0 else torch.tensor([shift], dtype=F64, device=op.device)

    def apply(v):
        return op.matvec_raw(None, v, sh)
    return apply


def _callable_apply(A, n: int, shift: float = 0.0):
    """A: scipy LinearOperator (numpy in/out, eig.py:97-101), torch callable on CUDA tensors, or NativeOperator.H."""
    rt = context()
    if hasattr(A, "matvec") and not isinstance(A, torch.Tensor):
        def base(v):
            return torch.from_numpy(np.asarray(A.matvec(v.cpu().numpy()), dtype=np.float64).reshape(-1)).to(rt.device)
    else:
        def base(v):
            with torch.no_grad():
                return dev_vec(A(v), rt.device)
    if shift == 0.0:
        return base
    return lambda v: _axpby(-shift, v, None, base(v))                      # (A - shift) v


def _triple(apply_A, apply_AT, n, k, which):
    lam, r = dominant_eigpair(apply_A, n, k, which)
    lam_l, l = dominant_eigpair(apply_AT, n, k, which)
    if abs(lam_l - lam) > 1e-8 * max(abs(lam), 1e-300):
        raise RuntimeError(f"left/right dominant eigenvalues disagree: {lam_l} vs {lam}")
    l = _axpby(0.0, l, 1.0 / _dot(l, r), l)                               # eig.py:36: l^T r = 1
    return lam, l, r


def _backward_vectors(apply_A_shifted, apply_AT_shifted, l, r, grad_l, grad_r):
    b = _axpby(-_dot(l, grad_l), r, None, grad_l.clone())                 # eig.py:53,139: grad_l - r (l . grad_l)
    lam_l0 = gmres_solve(apply_A_shifted, b)                              # :54,140
    b = _axpby(-_dot(r, grad_r), l, None, grad_r.clone())                 # :56,143
    lam_r0 = gmres_solve(apply_AT_shifted, b)                             # :57,144
    # Gauge.  (A - lambda) is singular: solutions differ by multiples of its null vector (r, resp. l).
    # scipy's GMRES from x0 = 0 stays inside the Krylov space of b, which lies in range(A - lambda) =
    # l-perp (resp. r-perp), so the reference implicitly returns the solution with l . x = 0 (r . x = 0).
    # We impose that condition explicitly so that it also holds after Krylov breakdown / restarts.
    lam_l0 = _axpby(-_dot(l, lam_l0), r, None, lam_l0)
    lam_r0 = _axpby(-_dot(r, lam_r0), l, None, lam_r0)
    return lam_l0, lam_r0


class DominantEig(torch.autograd.Function):
    """Dominant eigen-triple of a real diagonalisable matrix given as a torch.Tensor (eig.py:5-62).

    input:  A (n, n); k = Arnoldi basis size (ARPACK's ncv); which in {"LM", "SM", "LR", "SR"}.
    output: eigval (shape [1]), lefteigvector, righteigvector with l^T r = 1, r^T r = 1, on A's device.
    Only the gradient of A is computed (eig.py:61).
    """

    @staticmethod
    def forward(ctx, A, k, which="LM"):
        rt = context()
        op = DenseOperator(A)
        opT = DenseOperator(A.detach().t().contiguous())
        n = op.n_loc
        lam, l, r = _triple(_dense_apply(op), _dense_apply(opT), n, int(k), which)
        ctx.save_for_backward(A)
        ctx.lam, ctx.l, ctx.r = lam, l, r
        out_dev = A.device
        eigval = torch.tensor([lam], dtype=F64, device=out_dev)
        return eigval, l.to(out_dev), r.to(out_dev)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_eigval, grad_l, grad_r):
        A, = ctx.saved_tensors
        rt = context()
        lam, l, r = ctx.lam, ctx.l, ctx.r
        op = DenseOperator(A)
        opT = DenseOperator(A.detach().t().contiguous())
        gl, gr = dev_vec(grad_l, rt.device), dev_vec(grad_r, rt.device)
        lam_l0, lam_r0 = _backward_vectors(_dense_apply(op, lam), _dense_apply(opT, lam), l, r, gl, gr)
        ge = dev_vec(grad_eigval.reshape(-1), rt.device)
        # eig.py:58-60 as ONE rank-2 update: grad_A = l (ge r - lam_l0)^T - lam_r0 r^T  (dsea_outer + one fused add)
        n = l.numel()
        right = _axpby(ge, r, -1.0, lam_l0.clone())                       # ge r - lam_l0
        grad_A = torch.empty(n, n, dtype=F64, device=rt.device)
        _lib.check(rt.lib.dsea_outer(rt.handle, n, 1.0, ptr(l), ptr(right), grad_A.data_ptr(), stream_ptr()))
        grad_A.addr_(lam_r0, r, alpha=-1.0)
        return grad_A.to(A.device), None, None


DominantSparseEig = None


def setDominantSparseEig(A, AT, Aadjoint_to_gadjoint):
    """Creates the module-global `DominantSparseEig` primitive (eig.py:64-152).

    A, AT                  the matrix and its transpose as scipy LinearOperators (eig.py:97-101), or as
                           callables on CUDA tensors;
    Aadjoint_to_gadjoint   receives ((u1, v1), (u2, v2), (u3, v3)) — numpy arrays, adjoint of
                           A = sum_i u_i v_i^T — and returns the parameter adjoint as a torch.Tensor
                           (eig.py:103-110).
    """
    global DominantSparseEig

    def dim_of(x):
        if hasattr(x, "shape"):
            return int(x.shape[0])
        raise ValueError("A must expose .shape (scipy LinearOperator) so that its dimension is known")

    class _Primitive(torch.autograd.Function):
        @staticmethod
        def forward(ctx, g, k):
            n = dim_of(A)
            lam, l, r = _triple(_callable_apply(A, n), _callable_apply(AT, n), n, int(k), "LM")   # eig.py:116-117
            ctx.lam, ctx.l, ctx.r, ctx.n = lam, l, r, n
            out_dev = g.device if isinstance(g, torch.Tensor) else torch.device("cpu")
            return torch.tensor([lam], dtype=F64, device=out_dev), l.to(out_dev), r.to(out_dev)

        @staticmethod
        @torch.autograd.function.once_differentiable
        def backward(ctx, grad_eigval, grad_l, grad_r):
            rt = context()
            lam, l, r, n = ctx.lam, ctx.l, ctx.r, ctx.n
            gl, gr = dev_vec(grad_l, rt.device), dev_vec(grad_r, rt.device)
            lam_l0, lam_r0 = _backward_vectors(_callable_apply(A, n, lam), _callable_apply(AT, n, lam), l, r, gl, gr)
            ge = float(grad_eigval.reshape(-1)[0].item())
            npy = lambda t: t.detach().cpu().numpy()
            grad_A = ((ge * npy(l), npy(r)), (-npy(l), npy(lam_l0)), (-npy(lam_r0), npy(r)))     # eig.py:145-147
            return Aadjoint_to_gadjoint(grad_A), None

    _Primitive.__name__ = _Primitive.__qualname__ = "DominantSparseEig"
    DominantSparseEig = _Primitive
    return _Primitive
