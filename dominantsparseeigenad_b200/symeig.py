"""Dominant symmetric eigensolver primitives — same names as the reference's symeig.py.

    DominantSymeig.apply(A, k[, device])                                  (symeig.py:4-31)
    setDominantSparseSymeig(A, Aadjoint_to_padjoint)
    DominantSparseSymeig.apply(p, k, dim[, device])                       (symeig.py:33-88)

forward  = device-resident Lanczos (libdsea: matvec sweeps, fused two-pass re-orthogonalisation,
           single-CTA tridiagonal eigensolve, Ritz GEMV);
backward = projection + CG solve of (A - E0) x = b on the complement of psi0 + parameter-adjoint
           contraction, every piece itself a differentiable primitive so that
           torch.autograd.grad(..., create_graph=True) yields second and higher derivatives.
"""
from __future__ import annotations

import torch

from . import CG as _CG
from . import _lib
from .CG import _param_adjoint, _project_any
from .operators import as_operator, scale


class DominantSymeig(torch.autograd.Function):
    """Smallest eigenvalue / eigenvector of a dense real symmetric tensor A (symeig.py:4-31).

    input: A (n, n); k Lanczos vectors.  output: (eigval 0-dim, eigvector (n,)), on A's device.
    """

    @staticmethod
    def forward(ctx, A, k, device=torch.device("cpu")):
        op = as_operator(A, None, A.device)
        evals, vmin, _, _ = op.lanczos(None, int(k), _lib.DSEA_MIN)
        eigval, eigvector = evals[0].to(A.device), vmin.to(A.device)
        ctx.save_for_backward(A, eigval, eigvector)
        return eigval, eigvector

    @staticmethod
    def backward(ctx, grad_eigval, grad_eigvector):
        A, eigval, eigvector = ctx.saved_tensors
        Aprime = A - eigval * torch.eye(A.shape[0], device=A.device, dtype=A.dtype)     # symeig.py:25
        b = _project_any(eigvector, grad_eigvector)                                      # :27
        lambda0 = _CG.CGSubspace.apply(Aprime, b, eigvector)                             # :28
        grad_A = (scale(grad_eigval, eigvector) - lambda0)[:, None] * eigvector          # :29
        return grad_A, None, None


DominantSparseSymeig = None


def setDominantSparseSymeig(A, Aadjoint_to_gadjoint):
    """Creates the module-global `DominantSparseSymeig` primitive (symeig.py:33-88) and, like the
    reference, the companion `CG.CGSubspaceSparse` (symeig.py:67-69).

    A                     v -> A v: a Python callable or a native operator callable (`TFIM.H`, ...);
                          native operators run the whole solve inside libdsea.
    Aadjoint_to_gadjoint  (v1, v2) -> adjoint of the parameters given matrix adjoint v1 v2^T; must be
                          torch-differentiable for higher derivatives (README.md:88-126).
    """
    global DominantSparseSymeig
    _CG.setCGSubspaceSparse(A, Aadjoint_to_gadjoint)
    native = getattr(A, "_dsea_operator", None)
    state = {"op": native}

    def get_op(dim, device):
        op = state["op"]
        if op is None or (not getattr(op, "_dsea_native", False) and op.n_loc != dim):
            op = state["op"] = as_operator(A, dim, device, Aadjoint_to_gadjoint)
        return op

    class _Primitive(torch.autograd.Function):
        @staticmethod
        def forward(ctx, g, k, dim, device=torch.device("cpu")):
            call_dev = g.device if isinstance(g, torch.Tensor) else torch.device(device)
            op = get_op(int(dim), call_dev)
            evals, vmin, _, _ = op.lanczos(g, int(k), _lib.DSEA_MIN)                  # symeig.py:72-73
            eigval, eigvector = evals[0].to(call_dev), vmin.to(call_dev)
            ctx.op = op
            ctx.save_for_backward(g, eigval, eigvector)
            return eigval, eigvector

        @staticmethod
        def backward(ctx, grad_eigval, grad_eigvector):
            cg = _CG.CGSubspaceSparse.apply
            g, eigval, eigvector = ctx.saved_tensors
            b = _project_any(eigvector, grad_eigvector)                                # symeig.py:80
            lambda0 = cg(g, eigval, b, eigvector)                                      # :81
            v1, v2 = scale(grad_eigval, eigvector) - lambda0, eigvector                # :82-83
            grad_g = _param_adjoint(ctx.op, Aadjoint_to_gadjoint, v1, v2, g)           # :84
            return grad_g, None, None, None

    _Primitive.__name__ = _Primitive.__qualname__ = "DominantSparseSymeig"
    DominantSparseSymeig = _Primitive
    return _Primitive
