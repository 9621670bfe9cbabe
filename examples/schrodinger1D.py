"""Fit a 1-D potential so that the ground state matches a target wave function (config 1).

Same experiment as the reference's examples/schrodinger1D.py (N = 300, k = 300, LBFGS), with the three
forward variants running their eigen-solves on the GPU:
    --variant matrix    DominantSymeig on the dense Hamiltonian              (schrodinger1D.py:53-63)
    --variant callback  DominantSparseSymeig on the user's Python closures   (schrodinger1D.py:64-73, as shipped)
    --variant csr       DominantSparseSymeig on a native CSR + diag(V) operator (device resident)
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dominantsparseeigenad_b200 as dsea  # noqa: E402
import dominantsparseeigenad_b200.symeig as symeig  # noqa: E402


def fit(variant="csr", N=300, k=300, steps=5, verbose=True):
    """Runs `steps` L-BFGS steps (schrodinger1D.py:101-123) and returns the loss history."""
    import types
    args = types.SimpleNamespace(variant=variant, N=N, k=k, steps=steps)
    N, k = args.N, min(args.k, args.N)
    xmin, xmax = -1.0, 1.0
    x = np.linspace(xmin, xmax, num=N, endpoint=False)
    h = (xmax - xmin) / N
    t = np.zeros(N)
    idx = np.abs(x) < 0.5
    t[idx] = 1.0 - np.abs(x[idx])
    target = torch.from_numpy(t / np.linalg.norm(t))                                   # schrodinger1D.py:91-94
    dev = torch.device("cuda") if args.variant == "csr" else torch.device("cpu")
    target = target.to(dev)
    potential = torch.nn.Parameter((0.5 * torch.from_numpy(x) ** 2).to(dev))
    Kcsr = sp.diags([np.ones(N - 1), -2 * np.ones(N), np.ones(N - 1)], [-1, 0, 1], format="csr") * (-0.5 / h ** 2)
    Kdense = torch.from_numpy(Kcsr.toarray())
    op = dsea.SparseMatrixOperator.from_scipy(Kcsr, potential) if args.variant == "csr" else None

    def Hsparse(v):                                                                      # schrodinger1D.py:18-27
        zero = torch.zeros(1, dtype=torch.float64)
        return -0.5 / h ** 2 * (-2 * v + torch.cat((v[1:], zero)) + torch.cat((zero, v[:-1]))) + potential * v

    def forward():
        if args.variant == "matrix":
            _, psi0 = symeig.DominantSymeig.apply(Kdense + torch.diag(potential), k)
        elif args.variant == "callback":
            symeig.setDominantSparseSymeig(Hsparse, lambda v1, v2: v1 * v2)
            _, psi0 = symeig.DominantSparseSymeig.apply(potential, k, N)
        else:
            symeig.setDominantSparseSymeig(op.H, op.Hadjoint_to_padjoint)
            _, psi0 = symeig.DominantSparseSymeig.apply(potential, k, N, dev)
        return 1.0 - (psi0.abs() * target).sum()                                        # schrodinger1D.py:62

    optimizer = torch.optim.LBFGS([potential], max_iter=10, tolerance_change=1e-7, tolerance_grad=1e-7,
                                  line_search_fn="strong_wolfe")

    def closure():
        optimizer.zero_grad()
        loss = forward()
        loss.backward()
        return loss

    history = []
    for i in range(args.steps):
        t0 = time.time()
        loss = optimizer.step(closure)
        history.append(loss.item())
        if verbose:
            print(i, loss.item(), f"{time.time() - t0:.2f} s")
    return history


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="csr", choices=["matrix", "callback", "csr"])
    ap.add_argument("--N", type=int, default=300)
    ap.add_argument("--k", type=int, default=300)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    fit(a.variant, a.N, a.k, a.steps)


if __name__ == "__main__":
    main()
