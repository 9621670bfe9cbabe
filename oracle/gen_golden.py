"""Pins the oracle to the real reference and writes tests/golden/*.npz.   (TEST INFRASTRUCTURE)

Run ONLY in the build container, where the unmodified reference is mounted read-only:

    python -m oracle.gen_golden            # from the repo root

It imports /root/reference/DominantSparseEigenAD and /root/reference/examples/TFIM/TFIM.py with
three compatibility shims and no source edits:
  * torch.symeig  -> torch.linalg.eigh     (removed from torch >= 2.0; Lanczos.py:98)
  * matplotlib    -> stub module            (examples import it at module top)
  * torch.randn   -> deterministic replay   (so reference and oracle see identical start vectors)
then
  1. asserts the oracle restatement equals the reference (bit-exact integer tables; <=1e-12 on
     floating outputs with identical start vectors), and
  2. stores the REFERENCE's outputs as golden vectors.  The GPU box has no /root/reference, so the
     `-m gpu` tests compare the CUDA path with these files and with the oracle.
The upstream result files examples/TFIM/datas/*.npz are data, not source; they are copied verbatim
(as arrays) into tests/golden/upstream_tfim.npz.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DSEA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _install_shims():
    # torch >= 2.0 keeps `torch.symeig` only as a stub that raises; replace it unconditionally.
    torch.symeig = lambda A, eigenvectors=True: torch.linalg.eigh(A)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "examples", "TFIM"))
    sys.path.insert(0, os.path.join(REF, "examples"))


class _ReplayRandn:
    """Context manager: routes the reference's torch.randn(n, ...) calls to a draw callable."""

    def __init__(self, draw):
        self.draw = draw

    def __enter__(self):
        self._orig = torch.randn
        draw = self.draw

        def fake(*size, dtype=None, device=None, **kw):
            if len(size) == 1 and isinstance(size[0], int):
                v = draw(size[0])
                return v.to(dtype or torch.float64)
            return self._orig(*size, dtype=dtype, device=device, **kw)

        torch.randn = fake
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _eig_section(orc):
    # ------------------------------------------------------------------ 7. non-symmetric family (eig.py)
    import scipy.sparse.linalg as _sla
    _orig_gmres = _sla.gmres

    def _gmres_compat(A, b, *a, tol=None, **kw):          # scipy >= 1.12 renamed tol -> rtol (eig.py:54,57)
        if tol is not None:
            kw.setdefault("rtol", tol)
        return _orig_gmres(A, b, *a, **kw)

    _sla.gmres = _gmres_compat
    from DominantSparseEigenAD.eig import DominantEig as RefDominantEig
    eg = {}
    for tag, D, kk in (("D5", 5, 25), ("D8", 8, 40)):
        G = orc.mps_transfer_matrix(D, 2, 2024 + D)
        Gt = torch.from_numpy(G).requires_grad_(True)
        rng = np.random.default_rng(99 + D)
        a = float(rng.standard_normal())
        M = torch.from_numpy(rng.standard_normal((D * D, D * D)))
        lam, l, r = RefDominantEig.apply(Gt, kk)
        l0, r0 = l, r
        # gauge-invariant loss (invariant under r -> -r, l -> -l):  a*lambda + l^T M r   (test_gradient.py:16-19)
        loss = a * lam + l0.matmul(M).matmul(r0)
        gA, = torch.autograd.grad(loss, Gt)
        olam, ol, orr = orc.dominant_eig_triple(G, kk)
        l, r = l.detach(), r.detach()
        sgn = np.sign(np.dot(orr, r.numpy()))
        assert abs(olam[0] - lam.item()) < 1e-12 * abs(lam.item())
        assert np.allclose(sgn * orr, r.numpy(), atol=1e-10) and np.allclose(sgn * ol, l.numpy(), atol=1e-10)
        ogA = orc.dominant_eig_backward(G, lam.detach().numpy(), l.numpy(), r.numpy(), np.array([a]),
                                        (M @ r.numpy()), (M.numpy().T @ l.numpy()))
        assert np.allclose(ogA, gA.numpy(), atol=1e-9)
        eg.update({tag + "_G": G, tag + "_k": np.array(kk), tag + "_a": np.array(a), tag + "_M": M.numpy(),
                   tag + "_lam": lam.detach().numpy(), tag + "_l": l.numpy(), tag + "_r": r.numpy(),
                   tag + "_loss": np.array(loss.item()), tag + "_gradA": gA.numpy()})
        print("eig", tag, lam.item(), loss.item(), np.abs(gA.numpy()).max())
    np.savez_compressed(os.path.join(OUT, "dominant_eig.npz"), **eg)


def main():
    only_eig = "--only-eig" in sys.argv
    _install_shims()
    import contextlib
    import io

    import DominantSparseEigenAD.symeig as ref_symeig
    from DominantSparseEigenAD.CG import CG_torch as ref_cg
    from DominantSparseEigenAD.Lanczos import symeigLanczos as ref_symeig_lanczos
    from TFIM import TFIM as RefTFIM

    from oracle import dsea_oracle as orc

    os.makedirs(OUT, exist_ok=True)
    f32round = lambda x: float(np.float64(np.float32(x)))

    if only_eig:
        return _eig_section(orc)
    # ------------------------------------------------------------------ 1. integer tables
    tables = {}
    for N in (3, 4, 10, 12, 16):
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefTFIM(N)
        flips_ref = ref.flips_basis.numpy()
        diag_ref = ref.diag_elements.numpy()
        assert flips_ref.dtype == np.int64
        assert np.array_equal(flips_ref, orc.tfim_flip_table(N)), N
        assert np.array_equal(diag_ref, orc.tfim_diagonal(N)), N
        assert np.array_equal(diag_ref, orc.tfim_diagonal_closed_form(N, np.arange(1 << N))), N
        if N <= 10:
            tables[f"flips_N{N}"] = flips_ref
            tables[f"diag_N{N}"] = diag_ref
        tables[f"flips_sha256_N{N}"] = np.array(_sha(flips_ref))
        tables[f"diag_sha256_N{N}"] = np.array(_sha(diag_ref))
    np.savez_compressed(os.path.join(OUT, "tfim_tables.npz"), **tables)
    print("tables ok")

    # ------------------------------------------------------------------ 2. matvec / adjoint KATs
    kat = {}
    for N, g in ((4, 0.7), (10, 1.3), (12, 1.0)):
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefTFIM(N)
        ref.g = torch.tensor([g], dtype=torch.float64)
        mine = orc.TFIMOracle(N, g)
        rng = np.random.default_rng(1000 + N)
        v = torch.from_numpy(rng.standard_normal(1 << N))
        w = torch.from_numpy(rng.standard_normal(1 << N))
        Hv = ref.H(v)
        adj = ref.Hadjoint_to_gadjoint(w, v)
        assert torch.equal(Hv, mine.H(v))
        assert torch.equal(adj, mine.Hadjoint_to_gadjoint(w, v))
        kat[f"v_N{N}"], kat[f"w_N{N}"] = v.numpy(), w.numpy()
        kat[f"g_N{N}"] = np.array(g)
        kat[f"Hv_N{N}"], kat[f"adj_N{N}"] = Hv.numpy(), adj.numpy()
        kat[f"pHpg_v_N{N}"] = ref.pHpg(v).numpy()
    np.savez_compressed(os.path.join(OUT, "tfim_matvec_kat.npz"), **kat)
    print("matvec KATs ok")

    # ------------------------------------------------------------------ 3. Lanczos / CG KATs (dense, seeded)
    lk = {}
    rng = np.random.default_rng(7)
    n, k = 200, 120
    A = 0.1 * rng.random((n, n))
    A = torch.from_numpy(A + A.T)
    draws = orc.SeededDraws(11)
    q0, qp = draws(n), draws(n)
    with _ReplayRandn(orc.ListDraws([q0, qp])):
        rmin, rvmin, rmax, rvmax = ref_symeig_lanczos(A, k)
    omin, ovmin, omax, ovmax = orc.extreme_eigpair(lambda v: A.matmul(v), n, k, orc.ListDraws([q0, qp]), "both")
    assert abs(rmin - omin) < 1e-12 and abs(rmax - omax) < 1e-12
    assert torch.allclose(rvmin, ovmin, atol=1e-10) and torch.allclose(rvmax, ovmax, atol=1e-10)
    lk.update(A=A.numpy(), q0=q0.numpy(), k=np.array(k), eval_min=rmin.numpy(), evec_min=rvmin.numpy(),
              eval_max=rmax.numpy(), evec_max=rvmax.numpy())
    # low-rank CG (test_CG.py:29-47 shaped)
    n2 = 150
    B = rng.standard_normal((n2, n2))
    B = torch.from_numpy(B + B.T)
    w_, V_ = torch.linalg.eigh(B)
    psi = V_[:, 0]
    Bp = B - w_[0] * torch.eye(n2, dtype=torch.float64)
    b = torch.from_numpy(rng.standard_normal(n2))
    b = b - psi.dot(b) * psi
    x0 = torch.from_numpy(rng.standard_normal(n2))
    x0 = x0 - psi.dot(x0) * psi
    xr = ref_cg(Bp, b, x0)
    xo = orc.cg_solve(lambda v: Bp.matmul(v), b, x0)
    assert torch.allclose(xr, xo, atol=1e-12, rtol=0)
    lk.update(cg_A=Bp.numpy(), cg_b=b.numpy(), cg_x0=x0.numpy(), cg_psi=psi.numpy(), cg_x=xr.numpy())
    np.savez_compressed(os.path.join(OUT, "lanczos_cg_kat.npz"), **lk)
    print("Lanczos/CG KATs ok")

    # ------------------------------------------------------------------ 4. TFIM E0 / dE0 / d2E0 / chiF
    tf = {}
    cases = [(10, 300, [0.5, 1.0, 1.0050505050505052, 1.25, 1.5]), (12, 200, [1.0, 1.5])]
    for N, kk, gs in cases:
        with contextlib.redirect_stdout(io.StringIO()):
            ref = RefTFIM(N)
        for gi, g_raw in enumerate(gs):
            g = f32round(g_raw)                      # E0.py:95 builds g through a float32 tensor
            tag = f"N{N}_g{gi}"
            # --- E0 family (E0.py:53-67) with replayed start vectors
            seed = 5000 + 17 * N + gi
            ref.g = torch.tensor([g], dtype=torch.float64, requires_grad=True)
            with _ReplayRandn(orc.SeededDraws(seed)):
                ref_symeig.setDominantSparseSymeig(ref.H, ref.Hadjoint_to_gadjoint)
                E0, psi0 = ref_symeig.DominantSparseSymeig.apply(ref.g, kk, ref.dim)
                dE0, = torch.autograd.grad(E0, ref.g, create_graph=True)
                d2E0, = torch.autograd.grad(dE0, ref.g)
            mine = orc.TFIMOracle(N, torch.tensor([g], dtype=torch.float64, requires_grad=True))
            oE0, odE0, od2E0, opsi = orc.tfim_energy_derivatives(mine, kk, orc.SeededDraws(seed))
            assert abs(oE0 - E0.item()) < 1e-12 * abs(E0.item()), (oE0, E0.item())
            assert abs(odE0 - dE0.item()) < 1e-9 * abs(dE0.item()), (odE0, dE0.item())
            assert abs(od2E0 - d2E0.item()) < 1e-7 * abs(d2E0.item()), (od2E0, d2E0.item())
            assert abs(abs(opsi.dot(psi0.detach()).item()) - 1) < 1e-12
            # --- chiF (chiF.py:40-53)
            ref.g = torch.tensor([g], dtype=torch.float64, requires_grad=True)
            with _ReplayRandn(orc.SeededDraws(seed + 1)):
                ref_symeig.setDominantSparseSymeig(ref.H, ref.Hadjoint_to_gadjoint)
                E0c, psic = ref_symeig.DominantSparseSymeig.apply(ref.g, kk, ref.dim)
                logF = torch.log(psic.detach().matmul(psic))
                dlogF, = torch.autograd.grad(logF, ref.g, create_graph=True)
                d2logF, = torch.autograd.grad(dlogF, ref.g)
            chif = -d2logF.item()
            mine = orc.TFIMOracle(N, torch.tensor([g], dtype=torch.float64, requires_grad=True))
            _, _, ochif = orc.tfim_fidelity_susceptibility(mine, kk, orc.SeededDraws(seed + 1))
            assert abs(ochif - chif) < 1e-6 * abs(chif), (ochif, chif)
            an = orc.tfim_analytic(N, g)
            tf[tag + "_g"] = np.array(g)
            tf[tag + "_k"] = np.array(kk)
            tf[tag + "_seed"] = np.array(seed)
            tf[tag + "_ref"] = np.array([E0.item(), dE0.item(), d2E0.item(), chif])
            tf[tag + "_analytic"] = np.array(an)
            if N == 10:
                tf[tag + "_psi0"] = psi0.detach().numpy()
            print(tag, g, E0.item(), dE0.item(), d2E0.item(), chif, "analytic", an)
    np.savez_compressed(os.path.join(OUT, "tfim_derivatives.npz"), **tf)

    # ------------------------------------------------------------------ 5. upstream stored results
    up = {}
    for N in (10, 16, 20):
        e = np.load(os.path.join(REF, "examples", "TFIM", "datas", f"E0_N_{N}.npz"))
        c = np.load(os.path.join(REF, "examples", "TFIM", "datas", f"chiF_N_{N}.npz"))
        up[f"gs_N{N}"] = e["gs"]
        up[f"E0s_N{N}"], up[f"dE0s_N{N}"], up[f"d2E0s_N{N}"] = e["E0s"], e["dE0s"], e["d2E0s"]
        up[f"chiFs_N{N}"] = c["chiFs"]
        assert np.array_equal(e["gs"], c["gs"])
    np.savez_compressed(os.path.join(OUT, "upstream_tfim.npz"), **up)

    # ------------------------------------------------------------------ 6. config 1 (schrodinger1D as shipped)
    import schrodinger1D as ref_s1d
    N, kk = 300, 300
    xm = torch.from_numpy(np.linspace(-1.0, 1.0, num=N, endpoint=False))
    sch = {}
    for variant in ("matrixAD", "sparseAD"):
        model = ref_s1d.Schrodinger1D(-1.0, 1.0, N, xm)
        mine = orc.Schrodinger1DOracle(N)
        seed = 31337
        with _ReplayRandn(orc.SeededDraws(seed)):
            loss = getattr(model, "forward_" + variant)(mine.target, kk)
            grad, = torch.autograd.grad(loss, model.potential)
        oloss = (mine.loss_dense if variant == "matrixAD" else mine.loss_sparse)(kk, orc.SeededDraws(seed))
        ograd, = torch.autograd.grad(oloss, mine.potential)
        assert abs(oloss.item() - loss.item()) < 1e-10, (oloss.item(), loss.item())
        assert torch.allclose(ograd, grad, atol=1e-9, rtol=1e-6)
        sch[variant + "_loss"] = np.array(loss.item())
        sch[variant + "_grad"] = grad.numpy()
        print("config1", variant, loss.item(), grad.norm().item())
    sch["target"] = mine.target.numpy()
    np.savez_compressed(os.path.join(OUT, "schrodinger1d.npz"), **sch)
    _eig_section(orc)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
