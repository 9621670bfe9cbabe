"""Infinite-chain TFIM ground-state energy by variational uniform MPS — the consumer of the non-symmetric primitives.

Mirrors the role of examples/TFIM_vumps/general.py:44-100 in the reference: the energy per site of a
translation-invariant MPS with a general real tensor A[s, i, j] (physical d = 2, bond D) is

    e(A) = <l| (A A h A A) |r> / lambda^2,      T = sum_s A[s] (x) A[s]   (D^2 x D^2 transfer matrix),

with (lambda, l, r) the DOMINANT eigen-triple of T.  `DominantEig` (dense T on the GPU) and `DominantSparseEig`
(T and T^T as matrix-free closures, here CUDA einsum contractions of cost O(D^3) instead of O(D^4)) give
lambda, l, r with reverse-mode gradients; L-BFGS on A then converges to the exact energy
e0(g) = -(1 / 2 pi) int_{-pi}^{pi} sqrt(1 + g^2 - 2 g cos k) dk   (analytic.py:8-12 of the reference).

    python examples/tfim_vumps.py [--D 8] [--g 1.0] [--k 40] [--steps 30] [--sparse]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dominantsparseeigenad_b200 as dsea  # noqa: E402
from dominantsparseeigenad_b200 import eig  # noqa: E402

F64 = torch.float64


def exact_energy_per_site(g: float, nk: int = 200001) -> float:
    k = np.linspace(-math.pi, math.pi, nk)
    return float(-np.trapezoid(np.sqrt(1.0 + g * g - 2.0 * g * np.cos(k)), k) / (2.0 * math.pi))


def two_site_hamiltonian(g: float, device) -> torch.Tensor:
    """h[a, b, c, d] = <a b| -sz sz - (g/2)(sx 1 + 1 sx) |c d>."""
    sx = torch.tensor([[0.0, 1.0], [1.0, 0.0]], dtype=F64, device=device)
    sz = torch.tensor([[1.0, 0.0], [0.0, -1.0]], dtype=F64, device=device)
    one = torch.eye(2, dtype=F64, device=device)
    h = -torch.kron(sz, sz) - 0.5 * g * (torch.kron(sx, one) + torch.kron(one, sx))
    return h.reshape(2, 2, 2, 2)


class UniformMPS(torch.nn.Module):
    def __init__(self, D: int, g: float, k: int, device="cuda", seed: int = 0):
        super().__init__()
        self.D, self.k, self.g = D, k, g
        gen = torch.Generator().manual_seed(seed)
        self.A = torch.nn.Parameter(torch.randn(2, D, D, dtype=F64, generator=gen).to(device))
        self.h = two_site_hamiltonian(g, device)

    def _energy(self, lam, l, r):
        D, A = self.D, self.A
        l, r = l.reshape(D, D), r.reshape(D, D)
        up = torch.einsum("aik,bkj->abij", A, A)                    # two sites, upper layer
        lo = torch.einsum("cml,dln->cdmn", A, A)                    # lower layer
        return torch.einsum("abij,abcd,cdmn,im,jn->", up, self.h, lo, l, r) / lam.reshape(()) ** 2

    def energy_dense(self):
        D = self.D
        T = torch.einsum("kij,kmn->imjn", self.A, self.A).reshape(D * D, D * D)
        lam, l, r = eig.DominantEig.apply(T, self.k)               # general.py:50
        return self._energy(lam, l, r)

    def energy_matrix_free(self):
        D = self.D
        A = self.A.detach()

        class _Op:                                                   # "A is a function" with a known dimension
            shape = (D * D, D * D)

            def __init__(self, fn):
                self.fn = fn

            def __call__(self, v):
                return self.fn(v)

        right = _Op(lambda v: torch.einsum("kij,kmn,jn->im", A, A, v.reshape(D, D)).reshape(-1))
        left = _Op(lambda v: torch.einsum("kij,kmn,im->jn", A, A, v.reshape(D, D)).reshape(-1))

        def T_adjoint_to_A_adjoint(pairs):                           # general.py:69-77 (numpy pairs u v^T)
            grad = torch.zeros_like(A)
            for u, v in pairs:
                U = torch.as_tensor(u, dtype=F64, device=A.device).reshape(D, D)
                V = torch.as_tensor(v, dtype=F64, device=A.device).reshape(D, D)
                grad = grad + torch.einsum("im,jn,kmn->kij", U, V, A) + torch.einsum("mi,nj,kmn->kij", U, V, A)
            return grad

        eig.setDominantSparseEig(right, left, T_adjoint_to_A_adjoint)
        lam, l, r = eig.DominantSparseEig.apply(self.A, self.k)      # general.py:95-96
        return self._energy(lam, l, r)


def optimise(model: UniformMPS, steps: int, sparse: bool, verbose: bool = True):
    opt = torch.optim.LBFGS([model.A], max_iter=10, line_search_fn="strong_wolfe")
    forward = model.energy_matrix_free if sparse else model.energy_dense
    history = []

    def closure():
        opt.zero_grad()
        e = forward()
        e.backward()
        return e

    for it in range(steps):
        e = opt.step(closure)
        history.append(float(e.detach()))
        if verbose:
            print(f"step {it:3d}  e = {history[-1]:.12f}")
    return history


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=8)
    ap.add_argument("--g", type=float, default=1.0)
    ap.add_argument("--k", type=int, default=40)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--sparse", action="store_true")
    a = ap.parse_args()
    dsea.runtime.context()
    model = UniformMPS(a.D, a.g, min(a.k, a.D * a.D))
    hist = optimise(model, a.steps, a.sparse)
    e0 = exact_energy_per_site(a.g)
    print(f"D = {a.D}: e = {hist[-1]:.12f}   exact e0 = {e0:.12f}   error = {hist[-1] - e0:.3e}")


if __name__ == "__main__":
    main()
