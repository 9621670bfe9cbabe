"""Ground-state energy per site of the 1-D TFIM and its first two g-derivatives on the GPU.

Same experiment as the reference's examples/TFIM/E0.py (E0.py:70-113): a sweep over g with four methods —
analytic, torch full-spectrum AD (small N), DominantSymeig on the dense matrix (small N) and
DominantSparseSymeig on the matrix-free operator — but every eigen-solve runs in libdsea.

    python examples/tfim_E0.py --spins 16 --k 200 --points 11 [--save E0_N_16.npz]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dominantsparseeigenad_b200 as dsea  # noqa: E402
import dominantsparseeigenad_b200.symeig as symeig  # noqa: E402


def E0_analytic(N, g):
    """E0.py:9-23 (Neveu-Schwarz momenta; valid for every N, see oracle.tfim_analytic)."""
    m = torch.arange(N, dtype=torch.float64)
    ks = (2 * m + 1) * np.pi / N
    eps = 2 * torch.sqrt(g ** 2 - 2 * g * torch.cos(ks) + 1)
    E0 = -0.5 * eps.sum()
    dE0, = torch.autograd.grad(E0, g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, g)
    return E0.item() / N, dE0.item() / N, d2E0.item() / N


def E0_torchAD(model):                                        # E0.py:25-36
    Es, _ = torch.linalg.eigh(model.Hmatrix)
    dE0, = torch.autograd.grad(Es[0], model.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, model.g, retain_graph=True)
    return Es[0].item() / model.N, dE0.item() / model.N, d2E0.item() / model.N


def E0_matrixAD(model, k):                                    # E0.py:38-51
    E0, _ = symeig.DominantSymeig.apply(model.Hmatrix, k, model.device)
    dE0, = torch.autograd.grad(E0, model.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, model.g)
    return E0.item() / model.N, dE0.item() / model.N, d2E0.item() / model.N


def E0_sparseAD(model, k):                                    # E0.py:53-67
    symeig.setDominantSparseSymeig(model.H, model.Hadjoint_to_gadjoint)
    E0, _ = symeig.DominantSparseSymeig.apply(model.g, k, model.dim, model.device)
    dE0, = torch.autograd.grad(E0, model.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, model.g)
    return E0.item() / model.N, dE0.item() / model.N, d2E0.item() / model.N


def sweep(N, k, gs, save=""):
    """E0.py:94-113: the sparse-AD sweep over g; returns rows (g, E0/N, dE0/N, d2E0/N) and optionally writes the
    .npz file the reference's plot script reads (keys gs, E0s, dE0s, d2E0s)."""
    model = dsea.TFIM(N)
    k = min(k, model.dim)
    rows = []
    for gv in gs:
        model.g = torch.tensor([float(gv)], dtype=torch.float64, device=model.device, requires_grad=True)
        rows.append((float(gv),) + E0_sparseAD(model, k))
    if save:
        arr = np.array(rows)
        np.savez(save, gs=arr[:, 0], E0s=arr[:, 1], dE0s=arr[:, 2], d2E0s=arr[:, 3])
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spins", type=int, default=10)
    ap.add_argument("--k", type=int, default=300)
    ap.add_argument("--points", type=int, default=11)
    ap.add_argument("--save", default="")
    args = ap.parse_args()
    model = dsea.TFIM(args.spins)
    k = min(args.k, model.dim)
    gs = np.linspace(0.5, 1.5, num=args.points)
    out = np.empty((args.points, 3))
    dense_ok = args.spins <= 10
    print("g  E0/N (analytic, sparseAD[, torchAD, matrixAD])  dE0/N (...)  d2E0/N (...)")
    for i, gv in enumerate(gs):
        model.g = torch.tensor([gv], dtype=torch.float64, device=model.device, requires_grad=True)
        g_cpu = torch.tensor([gv], dtype=torch.float64, requires_grad=True)
        rows = [E0_analytic(args.spins, g_cpu), E0_sparseAD(model, k)]
        if dense_ok:
            model.setHmatrix()
            rows += [E0_torchAD(model), E0_matrixAD(model, k)]
        out[i] = rows[1]
        print(f"{gv:.4f}", *(" ".join(f"{r[j]:+.10f}" for r in rows) for j in range(3)), sep="  ")
    if args.save:
        np.savez(args.save, gs=gs, E0s=out[:, 0], dE0s=out[:, 1], d2E0s=out[:, 2])       # E0.py:111-113


if __name__ == "__main__":
    main()
