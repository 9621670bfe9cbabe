"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

Touches every kernel family once at CI sizes: TFIM sweeps (single tile, generic multi-sweep and the
persistent double-buffered kernel), fused reorth passes with ragged tiles, tridiagonal solver, CG, adjoint,
CSR / dense operators, Arnoldi.  Results are checked so a silent corruption also fails the run.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dominantsparseeigenad_b200 as dsea  # noqa: E402
from oracle import dsea_oracle as orc  # noqa: E402  (checker)


def main():
    rt = dsea.runtime.context()
    dev = rt.device
    rel = lambda a, b: abs(a - b) / abs(b)
    # TFIM, full E0 / dE0 / d2E0 with a multi-sweep plan (tiles of 2^6) ...
    rt.set_option("tfim_tile_bits", 6)
    N, k, g = 10, 60, 1.25
    m = dsea.TFIM(N)
    m.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
    E0, psi = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, dev)
    dE0, = torch.autograd.grad(E0, m.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, m.g)
    a = orc.tfim_analytic(N, g)
    assert rel(E0.item(), a[0]) < 1e-10 and rel(dE0.item(), a[1]) < 1e-6 and rel(d2E0.item(), a[2]) < 1e-6
    rt.set_option("tfim_tile_bits", 13)
    # ... and the persistent double-buffered kernel (needs >= 2 tiles of 2^13): one matvec + adjoint at N=15
    m2 = dsea.TFIM(15)
    m2.g = torch.tensor([0.9], dtype=torch.float64, device=dev)
    rng = np.random.default_rng(0)
    v, w = rng.standard_normal(1 << 15), rng.standard_normal(1 << 15)
    o = orc.TFIMOracle(15, 0.9)
    want = o.H(torch.from_numpy(v)).numpy()
    got = m2.H(torch.from_numpy(v).to(dev)).cpu().numpy()
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    adj = m2.Hadjoint_to_gadjoint(torch.from_numpy(w).to(dev), torch.from_numpy(v).to(dev)).item()
    assert rel(adj, o.Hadjoint_to_gadjoint(torch.from_numpy(w), torch.from_numpy(v)).item()) < 1e-11
    # round 2: fused normalisation in the pipelined first sweep (Lanczos at N=15), the fp32 shadow basis with its
    # Jacobi-Davidson polish (projected CG), and a plan with a DIRECT bit (L = 23: 13 + 9 shared-memory bits + 1)
    m2.g = torch.tensor([1.25], dtype=torch.float64, device=dev, requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(m2.H, m2.Hadjoint_to_gadjoint)
    E15, psi15 = dsea.symeig.DominantSparseSymeig.apply(m2.g, 160, m2.dim, dev)
    dE15, = torch.autograd.grad(E15, m2.g)
    dsea.runtime.set_basis_precision("fp32")
    m2.g = torch.tensor([1.25], dtype=torch.float64, device=dev, requires_grad=True)
    E15f, psi15f = dsea.symeig.DominantSparseSymeig.apply(m2.g, 160, m2.dim, dev)
    dsea.runtime.set_basis_precision("fp64")
    a15 = orc.tfim_analytic(15, 1.25)
    assert rel(E15f.item(), a15[0]) < 1e-10 and rel(E15.item(), a15[0]) < 1e-10 and rel(dE15.item(), a15[1]) < 1e-6
    from dominantsparseeigenad_b200 import selfcheck
    ident = selfcheck.operator_identities(dsea.TFIM(23))
    assert selfcheck.verdict(ident), ident
    # dense (odd n: ragged reorth tiles) forward + backward
    torch.manual_seed(0)
    n = 333
    K = torch.randn(n, n, dtype=torch.float64)
    K = K + K.T
    pot = torch.randn(n, dtype=torch.float64, requires_grad=True)
    H = K + torch.diag(pot)
    lam, vec = dsea.symeig.DominantSymeig.apply(H, 120)
    loss = vec.abs().sum() + lam
    gp, = torch.autograd.grad(loss, pot)
    w_, V_ = torch.linalg.eigh(H.detach())
    assert rel(lam.item(), w_[0].item()) < 1e-9 and torch.isfinite(gp).all()
    # CSR + diag, CG, Arnoldi
    import scipy.sparse as sp
    M = sp.random(257, 257, density=0.03, random_state=1, format="csr")
    M = (M + M.T).tocsr()
    op = dsea.SparseMatrixOperator.from_scipy(M, torch.zeros(257, dtype=torch.float64, device=dev))
    e, _ = dsea.Lanczos.symeigLanczos(op.H, 100, device=dev, extreme="min", sparse=True, dim=257)
    assert rel(e.item(), np.linalg.eigvalsh(M.toarray())[0]) < 1e-8
    from dominantsparseeigenad_b200.eig import DominantEig
    G = torch.from_numpy(orc.mps_transfer_matrix(4, 2, 3)).requires_grad_(True)
    lam, l, r = DominantEig.apply(G, 16)
    (lam.sum() + l.dot(r)).backward()
    assert torch.isfinite(G.grad).all()
    torch.cuda.synchronize()
    print("SANITIZE_WORKLOAD_OK launches", rt.launch_count())


if __name__ == "__main__":
    main()
