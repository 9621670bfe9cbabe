"""Conjugate-gradient solvers — same public names as the reference's DominantSparseEigenAD/CG.py.

    CG_torch(A, b, initialx, sparse=False)                     (CG.py:3-41)
    CGSubspace            autograd.Function, dense matrix      (CG.py:43-71)
    setCGSubspaceSparse(A, Aadjoint_to_gadjoint) -> module-global CGSubspaceSparse   (CG.py:73-140)

The iteration runs in libdsea: one operator application per iteration, fused x/r update with the
residual norm in its epilogue, device-resident scalars and convergence flag.  Both primitives stay
re-entrant (their backward calls the primitive again) so arbitrary-order derivatives work.
"""
from __future__ import annotations

import torch

from . import runtime
from .operators import Dot, as_operator, project, scale
from .runtime import context, dev_vec


def CG_torch(A, b, initialx, sparse=False):
    """Solves A x = b for SPD (or semidefinite-on-a-subspace) A; stops when |r| < 1e-7 (CG.py:25,35)."""
    out_dev = b.device
    if isinstance(A, torch.Tensor):
        op = as_operator(A, None, A.device)
    else:
        if not sparse and not hasattr(A, "_dsea_operator"):
            raise TypeError("A must be a torch.Tensor unless sparse=True (CG.py:18-23)")
        op = as_operator(A, b.shape[0], b.device)
    x = op.cg(None, None, b, initialx)
    return x.to(out_dev)


def _solve_on_subspace(op, param, E0, b, psi):
    """x0 random, projected off psi (CG.py:58-59,121-122); then CG on (A - E0).

    `runtime.cg_start = "zero"` (opt-in) starts from x0 = 0 instead.  The solution on the complement of psi is unique,
    so the result agrees to the CG tolerance, but a zero right-hand side — the backward pass of a loss that depends on
    E0 only, where the reference still runs ~100 iterations from its random start — then terminates at
    once.  The default stays the reference's random start so that benchmarks do the reference's work."""
    rt = context()
    if runtime.cg_start == "zero":
        x0 = torch.zeros(op.n_loc, dtype=torch.float64, device=rt.device)
    else:
        x0 = runtime.start_vector(op.n_loc, "cg")
        x0 = project(dev_vec(psi, rt.device), x0)
    return op.cg(param, E0, b, x0)


class CGSubspace(torch.autograd.Function):
    """x = A^+ b on the complement of alpha, A symmetric with alpha its zero mode (CG.py:43-71).

    input: A (n, n), b (n,), alpha (n,).  output: x with A x = b and alpha . x = 0.
    """

    @staticmethod
    def forward(ctx, A, b, alpha):
        op = as_operator(A, None, A.device)
        x = _solve_on_subspace(op, None, None, b, alpha).to(b.device)
        ctx.save_for_backward(A, alpha, x)
        return x

    @staticmethod
    def backward(ctx, grad_x):
        A, alpha, x = ctx.saved_tensors
        rhs = _project_any(alpha, grad_x)                       # CG.py:67
        grad_b = CGSubspace.apply(A, rhs, alpha)                # :68
        grad_A = -grad_b[:, None] * x                           # :69
        grad_alpha = -scale(_dot_any(alpha, grad_x), x)         # :70
        return grad_A, grad_b, grad_alpha


def _dot_any(a, b):
    """Differentiable dot that keeps the result on the inputs' device (CPU callers stay on CPU)."""
    out = Dot.apply(a, b)
    return out.to(a.device)


def _project_any(psi, v):
    return project(psi, v).to(v.device)


def _param_adjoint(op, user_adjoint, v1, v2, g):
    if getattr(op, "_dsea_native", False):
        out = op.adjoint(v1, v2, g)
    else:
        out = user_adjoint(v1, v2) if user_adjoint is not None else op.adjoint(v1, v2, g)
    return out.to(g.device) if isinstance(g, torch.Tensor) else out


CGSubspaceSparse = None


def setCGSubspaceSparse(A, Aadjoint_to_gadjoint):
    """Creates the module-global `CGSubspaceSparse` primitive for operator A (CG.py:73-140).

    `A` is a callable v -> A v (or a native operator callable such as `TFIM.H`);
    `Aadjoint_to_gadjoint(v1, v2)` maps the matrix adjoint v1 v2^T to the parameter adjoint.
    The operator dimension is taken from the right-hand side at call time.
    """
    global CGSubspaceSparse
    native = getattr(A, "_dsea_operator", None)

    class _Lazy:
        """Resolves callables to an operator once the vector length / device is known."""
        op = native

        @classmethod
        def get(cls, b):
            if cls.op is None or (not getattr(cls.op, "_dsea_native", False) and cls.op.n_loc != b.shape[0]):
                cls.op = as_operator(A, b.shape[0], b.device, Aadjoint_to_gadjoint)
            return cls.op

    class _Dispatch(torch.autograd.Function):
        @staticmethod
        def forward(ctx, g, E0, b, alpha):
            op = _Lazy.get(b)
            x = _solve_on_subspace(op, g, E0, b, alpha).to(b.device)
            ctx.g = g
            ctx.save_for_backward(E0, alpha, x)
            return x

        @staticmethod
        def backward(ctx, grad_x):
            g = ctx.g
            E0, alpha, x = ctx.saved_tensors
            op = _Lazy.get(x)
            rhs = _project_any(alpha, grad_x)
            grad_b = _Dispatch.apply(g, E0, rhs, alpha)        # CG.py:131,133 — bound to THIS operator, not the global
            v1, v2 = -grad_b, x
            grad_alpha = -scale(_dot_any(alpha, grad_x), x)
            grad_E0 = -_dot_any(v1, v2)
            grad_g = _param_adjoint(op, Aadjoint_to_gadjoint, v1, v2, g)
            return grad_g, grad_E0, grad_b, grad_alpha

    _Dispatch.__name__ = _Dispatch.__qualname__ = "CGSubspaceSparse"
    CGSubspaceSparse = _Dispatch
    return _Dispatch
