"""GPU parity: the CUDA path (through the ctypes C ABI) against the oracle on the same seeded inputs,
against the committed golden vectors (reference outputs), against the reference's own unit tests
re-hosted as known-answer tests, and — at sizes the oracle cannot reach — through size-independent
properties (symmetry, linearity, analytic values).

Tolerances (BASELINE.json north_star): integer maps bit-exact; eigenvalues 1e-10 relative;
eigenvector overlap >= 1 - 1e-8; gradients (dE0/dg, d2E0/dg2, chi_F) 1e-6 relative.
"""

import numpy as np
import pytest
import torch

from conftest import f32round

pytestmark = pytest.mark.gpu

F64 = torch.float64
EVAL_RTOL = 1e-10
OVERLAP_TOL = 1e-8
GRAD_RTOL = 1e-6


@pytest.fixture(scope="module")
def dsea():
    import dominantsparseeigenad_b200 as pkg
    pkg.runtime.context()           # raises loudly if libdsea.so / the GPU is missing
    yield pkg
    pkg.runtime.set_start_vector_hook(None)


@pytest.fixture(scope="module")
def orc():
    from oracle import dsea_oracle
    return dsea_oracle


def cuda(x):
    return torch.as_tensor(x, dtype=F64).cuda()


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


class ReferenceOrderDraws:
    """Feeds the CUDA path the start vectors the reference would draw from SeededDraws(seed):
    Lanczos consumes two draws (q0 and the discarded q', Lanczos.py:52,59), each CG solve one."""

    def __init__(self, orc, seed):
        self.draws = orc.SeededDraws(seed)

    def __call__(self, n, kind):
        v = self.draws(n)
        if kind == "lanczos":
            self.draws(n)
        return v


# ------------------------------------------------------------------------------------------------
# K1 / K6: TFIM matvec, dH/dg, adjoint
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [4, 10, 12])
def test_tfim_matvec_golden(dsea, golden, N):
    k = golden("tfim_matvec_kat.npz")
    m = dsea.TFIM(N)
    m.g = cuda([float(k[f"g_N{N}"])])
    v, w = cuda(k[f"v_N{N}"]), cuda(k[f"w_N{N}"])
    scale = np.abs(k[f"Hv_N{N}"]).max()
    assert np.abs(m.H(v).cpu().numpy() - k[f"Hv_N{N}"]).max() <= 1e-14 * scale * N
    assert np.abs(m.pHpg(v).cpu().numpy() - k[f"pHpg_v_N{N}"]).max() <= 1e-14 * scale * N
    adj = m.Hadjoint_to_gadjoint(w, v)
    assert adj.shape == (1,)
    assert rel(adj.item(), float(k[f"adj_N{N}"][0])) < 1e-12


@pytest.mark.parametrize("N,tile_bits,run_bits", [(12, 13, 0), (12, 5, 0), (12, 6, 2), (14, 8, 3), (16, 13, 0),
                                                  (16, 9, 0), (17, 7, 2), (18, 13, 4)])
def test_tfim_matvec_every_sweep_plan_vs_oracle(dsea, orc, N, tile_bits, run_bits):
    """Small tiles force the multi-sweep (strided high-bit) code paths that N >= 24 uses in production."""
    rt = dsea.runtime.context()
    rt.set_option("tfim_tile_bits", tile_bits)
    rt.set_option("tfim_run_bits", run_bits)
    try:
        g = 0.8 + 0.01 * N
        o = orc.TFIMOracle(N, g)
        rng = np.random.default_rng(N * 100 + tile_bits)
        v, w = rng.standard_normal(1 << N), rng.standard_normal(1 << N)
        m = dsea.TFIM(N)
        m.g = cuda([g])
        want = o.H(torch.from_numpy(v)).numpy()
        got = m.H(cuda(v)).cpu().numpy()
        assert np.abs(got - want).max() <= 1e-14 * N * np.abs(want).max()
        want_adj = o.Hadjoint_to_gadjoint(torch.from_numpy(w), torch.from_numpy(v)).item()
        assert rel(m.Hadjoint_to_gadjoint(cuda(w), cuda(v)).item(), want_adj) < 1e-11
        assert np.abs(m.pHpg(cuda(v)).cpu().numpy() - o.pHpg(torch.from_numpy(v)).numpy()).max() <= 1e-13 * N * np.abs(v).max()
    finally:
        rt.set_option("tfim_tile_bits", 13)
        rt.set_option("tfim_run_bits", 0)


def test_tfim_device_diagonal_bit_exact(dsea, orc):
    """H(e_s) with g = 0 returns diag[s] e_s: the device diagonal must equal TFIM.py:39-46 exactly."""
    N = 12
    m = dsea.TFIM(N)
    m.g = cuda([0.0])
    ones = torch.ones(1 << N, dtype=F64).cuda()
    assert np.array_equal(m.H(ones).cpu().numpy(), orc.tfim_diagonal(N))


def test_tfim_flip_map_bit_exact_on_device(dsea, orc):
    """(dH/dg) applied to basis-labelled vectors reproduces the flip table: -sum_i x[s ^ (1<<i)] with
    x[s] = 2^-s-style weights is exact in fp64 for N <= 10, so any wrong index shows up bit-for-bit."""
    N = 10
    m = dsea.TFIM(N)
    x = np.arange(1 << N, dtype=np.float64)          # integers: sums are exact
    got = m.pHpg(cuda(x)).cpu().numpy()
    want = -x[orc.tfim_flip_table(N)].sum(axis=1)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("N", [20, 22])
def test_tfim_properties_at_scale(dsea, N):
    """Sizes beyond the oracle's comfort: symmetry, linearity, adjoint identity."""
    m = dsea.TFIM(N)
    m.g = cuda([1.1])
    gen = torch.Generator(device="cuda").manual_seed(N)
    v = torch.randn(1 << N, dtype=F64, device="cuda", generator=gen)
    w = torch.randn(1 << N, dtype=F64, device="cuda", generator=gen)
    Hv, Hw = m.H(v), m.H(w)
    assert rel(torch.dot(w, Hv).item(), torch.dot(Hw, v).item()) < 1e-11             # symmetric
    assert (m.H(2.0 * v - 3.0 * w) - (2.0 * Hv - 3.0 * Hw)).abs().max().item() < 1e-10 * Hv.abs().max().item()
    m0 = dsea.TFIM(N)
    m0.g = cuda([0.1])
    dH = (Hv - m0.H(v)) / (1.1 - 0.1)                                                  # H is affine in g
    assert (dH - m.pHpg(v)).abs().max().item() < 1e-10 * Hv.abs().max().item()
    assert rel(m.Hadjoint_to_gadjoint(w, v).item(), torch.dot(w, m.pHpg(v)).item()) < 1e-11
    # uniform state: (dH/dg) 1 = -N 1 exactly
    ones = torch.ones(1 << N, dtype=F64, device="cuda")
    assert torch.equal(m.pHpg(ones), -N * ones)


# ------------------------------------------------------------------------------------------------
# K2-K5: Lanczos
# ------------------------------------------------------------------------------------------------
def test_lanczos_dense_golden(dsea, golden):
    k = golden("lanczos_cg_kat.npz")
    A, q0 = cuda(k["A"]), torch.from_numpy(k["q0"])
    dsea.runtime.set_start_vector_hook(lambda n, kind: q0)
    try:
        lo, vlo, hi, vhi = dsea.Lanczos.symeigLanczos(A, int(k["k"]), extreme="both")
    finally:
        dsea.runtime.set_start_vector_hook(None)
    assert lo.dim() == 0 and vlo.shape == (A.shape[0],)
    assert rel(lo.item(), float(k["eval_min"])) < EVAL_RTOL and rel(hi.item(), float(k["eval_max"])) < EVAL_RTOL
    assert 1 - abs(np.dot(vlo.cpu().numpy(), k["evec_min"])) < OVERLAP_TOL
    assert 1 - abs(np.dot(vhi.cpu().numpy(), k["evec_max"])) < OVERLAP_TOL


@pytest.mark.parametrize("sparse", [False, True])
def test_lanczos_reference_test_normal(dsea, sparse):
    """test_Lanczos.py:6-31 / :64-78 re-hosted: n=1000, k=300, A = 0.1 rand + sym."""
    torch.manual_seed(0)
    n, k = 1000, 300
    A = 0.1 * torch.rand(n, n, dtype=F64, device="cuda")
    A = A + A.T
    if sparse:
        out = dsea.Lanczos.symeigLanczos(lambda v: torch.matmul(A, v), k, device=torch.device("cuda"),
                                         sparse=True, dim=n)
    else:
        out = dsea.Lanczos.symeigLanczos(A, k, device=torch.device("cuda"))
    emin, vmin, emax, vmax = out
    w, V = torch.linalg.eigh(A)
    assert torch.allclose(emin, w[0]) and torch.allclose(emax, w[-1])
    assert rel(emin.item(), w[0].item()) < EVAL_RTOL and rel(emax.item(), w[-1].item()) < EVAL_RTOL
    assert torch.allclose(vmin, V[:, 0]) or torch.allclose(vmin, -V[:, 0])
    assert torch.allclose(vmax, V[:, -1]) or torch.allclose(vmax, -V[:, -1])


def test_lanczos_reference_test_tridiagonal_k_equals_n(dsea):
    """test_Lanczos.py:100-127 re-hosted: FD Hamiltonian, k = n = 1000 (needs the breakdown guard)."""
    N = 1000
    x = torch.from_numpy(np.linspace(-1.0, 1.0, num=N, endpoint=False)).cuda()
    h = 2.0 / N
    K = -0.5 / h ** 2 * (torch.diag(-2 * torch.ones(N, dtype=F64)) + torch.diag(torch.ones(N - 1, dtype=F64), 1)
                         + torch.diag(torch.ones(N - 1, dtype=F64), -1)).cuda()
    H = K + torch.diag(0.5 * x ** 2)
    E0, psi0 = dsea.Lanczos.symeigLanczos(H, N, extreme="min")
    Es, psis = torch.linalg.eigh(H)
    assert torch.allclose(E0, Es[0])
    assert torch.allclose(psi0, psis[:, 0]) or torch.allclose(psi0, -psis[:, 0])


def test_lanczos_returns_orthonormal_basis_and_T(dsea):
    torch.manual_seed(1)
    n, k = 777, 60          # odd n exercises the unaligned tail paths
    A = torch.randn(n, n, dtype=F64, device="cuda")
    A = A + A.T
    Qk, T = dsea.Lanczos.Lanczos(A, k, device=torch.device("cuda"))
    assert Qk.shape == (n, k) and T.shape == (k, k)
    assert (Qk.T @ Qk - torch.eye(k, dtype=F64, device="cuda")).abs().max().item() < 1e-12
    assert (Qk.T @ A @ Qk - T).abs().max().item() < 1e-9 * A.abs().max().item() * n ** 0.5


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1024, 1025, 4099])
def test_level1_ragged_sizes(dsea, n):
    gen = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn(n, dtype=F64, device="cuda", generator=gen)
    b = torch.randn(n, dtype=F64, device="cuda", generator=gen)
    assert rel(dsea.dot(a, b).item(), torch.dot(a, b).item()) < 1e-12 or abs(torch.dot(a, b).item()) < 1e-12
    psi = a / a.norm()
    want = b - torch.dot(psi, b) * psi
    assert (dsea.project(psi, b) - want).abs().max().item() < 1e-13 * (1 + b.abs().max().item())


# ------------------------------------------------------------------------------------------------
# K7 / K8: CG
# ------------------------------------------------------------------------------------------------
def test_cg_golden_lowrank(dsea, golden):
    k = golden("lanczos_cg_kat.npz")
    x = dsea.CG.CG_torch(cuda(k["cg_A"]), cuda(k["cg_b"]), cuda(k["cg_x0"]))
    # both stop at |r| < 1e-7; different iteration counts (1 vs 2 matvecs/iter) => compare to that level
    A = k["cg_A"]
    assert np.linalg.norm(A @ x.cpu().numpy() - k["cg_b"]) < 1e-6
    assert abs(np.dot(x.cpu().numpy(), k["cg_psi"])) < 1e-6
    lam = np.linalg.eigvalsh(A)
    assert np.linalg.norm(x.cpu().numpy() - k["cg_x"]) < 4e-7 / lam[1]


def test_cg_reference_test_fullrank(dsea):
    """test_CG.py:5-27 / :51-70 re-hosted."""
    from scipy.stats import ortho_group
    rng = np.random.default_rng(5)
    n = 100
    U = ortho_group.rvs(n, random_state=5)
    A = cuda(U @ np.diag(1.0 + 10.0 * rng.random(n)) @ U.T)
    b, x0 = cuda(rng.standard_normal(n)), cuda(rng.standard_normal(n))
    x = dsea.CG.CG_torch(A, b, x0)
    assert torch.allclose(A.matmul(x), b)
    assert torch.allclose(x, torch.linalg.solve(A, b))


@pytest.mark.parametrize("sparse", [False, True])
def test_cg_reference_test_lowrank(dsea, sparse):
    """test_CG.py:29-47 re-hosted: A - lambda0 I, b and x0 orthogonal to the zero mode."""
    torch.manual_seed(2)
    n = 300
    A = torch.randn(n, n, dtype=F64, device="cuda")
    A = A + A.T
    w, V = torch.linalg.eigh(A)
    psi = V[:, 0]
    Ap = A - w[0] * torch.eye(n, dtype=F64, device="cuda")
    b = torch.randn(n, dtype=F64, device="cuda")
    b = b - torch.matmul(psi, b) * psi
    x0 = torch.randn(n, dtype=F64, device="cuda")
    x0 = x0 - torch.matmul(psi, x0) * psi
    x = dsea.CG.CG_torch((lambda v: Ap.matmul(v)) if sparse else Ap, b, x0, sparse=sparse)
    assert torch.allclose(Ap.matmul(x) - b, torch.zeros(n, dtype=F64, device="cuda"), atol=1e-6)
    assert abs(torch.matmul(x, psi).item()) < 1e-6


def test_cg_subspace_backward_analytic_2x2(dsea):
    """demos/CG_backward.py:32-44: closed-form derivative of the low-rank solve on a 2x2 system."""
    torch.manual_seed(3)
    A0 = torch.randn(2, 2, dtype=F64)
    A0 = A0 + A0.T
    b0 = torch.randn(2, dtype=F64)
    flip = torch.tensor([[0.0, -1.0], [1.0, 0.0]], dtype=F64)
    while True:
        alpha0 = torch.randn(2, dtype=F64)
        alpha0 = alpha0 / alpha0.norm()
        alpha = flip.matmul(alpha0)
        if alpha.matmul(A0).matmul(alpha) > 0:
            break
    alpha0.requires_grad_(True)
    alpha = flip.matmul(alpha0)
    z = torch.randn(2, dtype=F64)
    A = alpha.matmul(A0).matmul(alpha) * (alpha[:, None] * alpha)
    b = torch.matmul(alpha, b0) * alpha
    x = dsea.CG.CGSubspace.apply(A, b, alpha0)
    d, = torch.autograd.grad(torch.matmul(x, z), alpha0)
    xa = torch.matmul(alpha, b0) * alpha / torch.matmul(alpha, alpha) / alpha.matmul(A0).matmul(alpha)
    da = torch.matmul(xa, z) * (b0 / torch.matmul(alpha, b0) + z / torch.matmul(alpha, z)
                                - 2 * alpha / torch.matmul(alpha, alpha)
                                - 2 * torch.matmul(A0, alpha) / alpha.matmul(A0).matmul(alpha))
    da = flip.T.matmul(da)
    assert torch.allclose(x, xa.detach(), atol=1e-6)
    assert torch.allclose(d, da.detach(), atol=1e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------------
# primitives: dense (config 1 family)
# ------------------------------------------------------------------------------------------------
def test_dominant_symeig_reference_test(dsea):
    """test_symeig.py:5-46 re-hosted (CPU tensors in, CPU tensors out, compute on the GPU)."""
    torch.manual_seed(4)
    N = 300
    K = torch.randn(N, N, dtype=F64)
    K = K + K.T
    target = torch.randn(N, dtype=F64)
    potential = torch.randn(N, dtype=F64, requires_grad=True)
    H = K + torch.diag(potential)
    Es, psis = torch.linalg.eigh(H)
    loss_t = 1.0 - torch.matmul(psis[:, 0], target)
    grad_t, = torch.autograd.grad(loss_t, potential)
    _, psi0 = dsea.symeig.DominantSymeig.apply(H, 300)
    assert psi0.device.type == "cpu"
    loss_d = 1.0 - torch.matmul(psi0, target)
    grad_d, = torch.autograd.grad(loss_d, potential)
    assert torch.allclose(loss_d, loss_t) or torch.allclose(loss_d, 2.0 - loss_t)
    assert torch.allclose(grad_d, grad_t) or torch.allclose(grad_d, -grad_t)


def _schrodinger_parts(orc, device):
    mdl = orc.Schrodinger1DOracle(300)
    K = mdl.kinetic_dense()
    return mdl, K.to(device), mdl.target.to(device)


def test_config1_schrodinger_dense_golden(dsea, orc, golden):
    s = golden("schrodinger1d.npz")
    mdl, K, target = _schrodinger_parts(orc, "cpu")
    H = K + torch.diag(mdl.potential)
    _, psi0 = dsea.symeig.DominantSymeig.apply(H, 300)
    loss = 1.0 - (psi0.abs() * target).sum()
    grad, = torch.autograd.grad(loss, mdl.potential)
    assert abs(loss.item() - float(s["matrixAD_loss"])) < 1e-9
    assert np.allclose(grad.numpy(), s["matrixAD_grad"], rtol=GRAD_RTOL, atol=1e-6 * np.abs(s["matrixAD_grad"]).max())


def test_config1_schrodinger_cpu_callback_as_shipped(dsea, orc, golden):
    """schrodinger1D.py:64-73: user closures on CPU tensors; the solver stages vectors for A(v) only."""
    s = golden("schrodinger1d.npz")
    mdl, _, target = _schrodinger_parts(orc, "cpu")
    dsea.symeig.setDominantSparseSymeig(mdl.Hsparse, mdl.Hadjoint_to_padjoint)
    _, psi0 = dsea.symeig.DominantSparseSymeig.apply(mdl.potential, 300, 300)
    loss = 1.0 - (psi0.abs() * target).sum()
    grad, = torch.autograd.grad(loss, mdl.potential)
    assert abs(loss.item() - float(s["sparseAD_loss"])) < 1e-9
    assert np.allclose(grad.numpy(), s["sparseAD_grad"], rtol=GRAD_RTOL, atol=1e-6 * np.abs(s["sparseAD_grad"]).max())


def test_config1_schrodinger_native_csr(dsea, orc, golden):
    """Same Hamiltonian as an explicit CSR matrix + trainable diagonal, fully device resident."""
    import scipy.sparse as sp
    s = golden("schrodinger1d.npz")
    mdl, _, target = _schrodinger_parts(orc, "cuda")
    N, h = 300, 2.0 / 300
    Kcsr = sp.diags([np.ones(N - 1), -2 * np.ones(N), np.ones(N - 1)], [-1, 0, 1], format="csr") * (-0.5 / h ** 2)
    pot = mdl.potential.detach().cuda().requires_grad_(True)
    op = dsea.SparseMatrixOperator.from_scipy(Kcsr, pot)
    dsea.symeig.setDominantSparseSymeig(op.H, op.Hadjoint_to_padjoint)
    _, psi0 = dsea.symeig.DominantSparseSymeig.apply(pot, 300, 300, torch.device("cuda"))
    loss = 1.0 - (psi0.abs() * target).sum()
    grad, = torch.autograd.grad(loss, pot)
    assert abs(loss.item() - float(s["sparseAD_loss"])) < 1e-9
    assert np.allclose(grad.cpu().numpy(), s["sparseAD_grad"], rtol=GRAD_RTOL, atol=1e-6 * np.abs(s["sparseAD_grad"]).max())


# ------------------------------------------------------------------------------------------------
# primitives: TFIM (configs 2-3 family)
# ------------------------------------------------------------------------------------------------
def _tfim_E0_family(dsea, N, g, k):
    m = dsea.TFIM(N)
    m.g = torch.tensor([g], dtype=F64, device="cuda", requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
    E0, psi0 = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, torch.device("cuda"))      # E0.py:60-62
    dE0, = torch.autograd.grad(E0, m.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, m.g)
    return E0.item(), dE0.item(), d2E0.item(), psi0.detach()


def _tfim_chif(dsea, N, g, k):
    m = dsea.TFIM(N)
    m.g = torch.tensor([g], dtype=F64, device="cuda", requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
    E0, psi0 = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, torch.device("cuda"))      # chiF.py:46-52
    logF = torch.log(dsea.dot(psi0.detach(), psi0))
    dlogF, = torch.autograd.grad(logF, m.g, create_graph=True)
    d2logF, = torch.autograd.grad(dlogF, m.g)
    return -d2logF.item()


@pytest.mark.parametrize("tag", ["N10_g0", "N10_g1", "N10_g2", "N10_g3", "N10_g4", "N12_g0", "N12_g1"])
def test_tfim_derivatives_vs_reference_golden(dsea, orc, golden, tag):
    """Same g, k and start vectors as the reference run that produced the golden file."""
    d = golden("tfim_derivatives.npz")
    N = int(tag[1:3])
    g, k, seed = float(d[tag + "_g"]), int(d[tag + "_k"]), int(d[tag + "_seed"])
    ref = d[tag + "_ref"]
    dsea.runtime.set_start_vector_hook(ReferenceOrderDraws(orc, seed))
    try:
        E0, dE0, d2E0, psi0 = _tfim_E0_family(dsea, N, g, k)
        dsea.runtime.set_start_vector_hook(ReferenceOrderDraws(orc, seed + 1))
        chif = _tfim_chif(dsea, N, g, k)
    finally:
        dsea.runtime.set_start_vector_hook(None)
    assert rel(E0, ref[0]) < EVAL_RTOL
    assert rel(dE0, ref[1]) < GRAD_RTOL
    if g >= 1.0:        # below g=1 the reference itself is unstable at the 1e-4 level
        assert rel(d2E0, ref[2]) < GRAD_RTOL
        assert rel(chif, ref[3]) < GRAD_RTOL
    if (tag + "_psi0") in d.files:
        assert 1 - abs(np.dot(psi0.cpu().numpy(), d[tag + "_psi0"])) < OVERLAP_TOL


@pytest.mark.parametrize("N,k", [(10, 300), (16, 200)])
def test_tfim_vs_upstream_result_files(dsea, golden, N, k):
    """examples/TFIM/datas/E0_N_*.npz, chiF_N_*.npz (per-site values; g rounded through float32)."""
    up = golden("upstream_tfim.npz")
    for i in (50, 75, 99):
        g = f32round(up[f"gs_N{N}"][i])
        E0, dE0, d2E0, _ = _tfim_E0_family(dsea, N, g, k)
        assert rel(E0 / N, up[f"E0s_N{N}"][i]) < EVAL_RTOL
        assert rel(dE0 / N, up[f"dE0s_N{N}"][i]) < GRAD_RTOL
        assert rel(d2E0 / N, up[f"d2E0s_N{N}"][i]) < GRAD_RTOL
        assert rel(_tfim_chif(dsea, N, g, k), up[f"chiFs_N{N}"][i]) < GRAD_RTOL


def test_config2_tfim_N20_k100_vs_analytic_and_upstream(dsea, orc, golden):
    """BASELINE config 2: N=20, k=100, E0 / dE0 / d2E0 on one B200."""
    up = golden("upstream_tfim.npz")
    N, k, i = 20, 100, 50
    g = f32round(up["gs_N20"][i])
    E0, dE0, d2E0, psi0 = _tfim_E0_family(dsea, N, g, k)
    aE0, adE0, ad2E0, achi = orc.tfim_analytic(N, g)
    assert rel(E0, aE0) < EVAL_RTOL and rel(E0 / N, up["E0s_N20"][i]) < EVAL_RTOL
    assert rel(dE0, adE0) < GRAD_RTOL and rel(dE0 / N, up["dE0s_N20"][i]) < GRAD_RTOL
    assert rel(d2E0, ad2E0) < GRAD_RTOL and rel(d2E0 / N, up["d2E0s_N20"][i]) < GRAD_RTOL
    assert abs(psi0.norm().item() - 1.0) < 1e-12
    chif = _tfim_chif(dsea, N, g, k)
    assert rel(chif, achi) < GRAD_RTOL and rel(chif, up["chiFs_N20"][i]) < GRAD_RTOL


def test_tfim_four_methods_agree_N8(dsea, orc):
    """E0.py:94-108 shape: analytic, torch full-spectrum AD, DominantSymeig (dense) and DominantSparseSymeig
    (matrix-free) give the same E0, dE0/dg and d2E0/dg2; chi_F also matches the perturbation formula
    (chiF.py:11-24)."""
    N, k, g = 8, 256, 1.2
    m = dsea.TFIM(N)
    m.g = torch.tensor([g], dtype=F64, device="cuda", requires_grad=True)
    aE0, adE0, ad2E0, achi = orc.tfim_analytic(N, g)
    # matrix-free
    dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
    E0, _ = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, torch.device("cuda"))
    dE0, = torch.autograd.grad(E0, m.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, m.g)
    # dense primitive and torch's own full-spectrum AD on the same (device-built) matrix
    Hm = m.setHmatrix()
    assert (Hm.detach().cpu() - orc.TFIMOracle(N, g).dense()).abs().max().item() < 1e-10
    E0m, _ = dsea.symeig.DominantSymeig.apply(Hm, k, torch.device("cuda"))
    dE0m, = torch.autograd.grad(E0m, m.g, create_graph=True)
    d2E0m, = torch.autograd.grad(dE0m, m.g)
    Es, psis = torch.linalg.eigh(m.setHmatrix())
    dE0t, = torch.autograd.grad(Es[0], m.g, create_graph=True)
    d2E0t, = torch.autograd.grad(dE0t, m.g)
    for got in ((E0, dE0, d2E0), (E0m, dE0m, d2E0m), (Es[0], dE0t, d2E0t)):
        assert rel(got[0].item(), aE0) < EVAL_RTOL
        assert rel(got[1].item(), adE0) < GRAD_RTOL
        assert rel(got[2].item(), ad2E0) < 1e-5          # the dense variants carry the reference's 1e-12 noise
    # chi_F by the full-spectrum perturbation formula
    P = m.setpHpg()
    num = (psis[:, 0].detach() @ P @ psis.detach())[1:] ** 2
    chi_pert = (num / (Es[0] - Es[1:]).detach() ** 2).sum().item()
    assert rel(chi_pert, achi) < 1e-6 and rel(_tfim_chif(dsea, N, g, k), achi) < GRAD_RTOL


def test_eigenvector_residual_N18(dsea):
    """|H psi0 - E0 psi0| small and E0 matches the analytic value where no CPU oracle is cheap."""
    from oracle import dsea_oracle as orc
    N, k, g = 18, 120, 1.25
    m = dsea.TFIM(N)
    m.g = cuda([g])
    E0, psi0 = dsea.Lanczos.symeigLanczos(m.H, k, device=torch.device("cuda"), extreme="min", sparse=True, dim=m.dim)
    assert rel(E0.item(), orc.tfim_analytic(N, g)[0]) < EVAL_RTOL
    assert (m.H(psi0) - E0 * psi0).norm().item() < 1e-7


def test_device_randn_is_standard_normal(dsea):
    n = 1 << 20
    a = dsea.runtime.start_vector(n, "lanczos")
    b = dsea.runtime.start_vector(n, "lanczos")
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1.0) < 5e-3
    assert abs((a ** 4).mean().item() - 3.0) < 0.05
    assert abs(torch.dot(a, b).item()) / n < 5e-3           # successive draws are independent
    torch.manual_seed(1234)
    dsea.runtime._draw_counter = 0
    c = dsea.runtime.start_vector(1000, "cg")
    torch.manual_seed(1234)
    dsea.runtime._draw_counter = 0
    assert torch.equal(c, dsea.runtime.start_vector(1000, "cg"))


# ------------------------------------------------------------------------------------------------
# non-symmetric family (eig.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["D5", "D8"])
def test_dominant_eig_golden(dsea, golden, tag):
    """Forward triple and grad_A against the reference's own output (ARPACK + GMRES on the CPU)."""
    from dominantsparseeigenad_b200.eig import DominantEig
    e = golden("dominant_eig.npz")
    G = torch.from_numpy(e[tag + "_G"]).requires_grad_(True)
    k, a, M = int(e[tag + "_k"]), float(e[tag + "_a"]), torch.from_numpy(e[tag + "_M"])
    lam, l, r = DominantEig.apply(G, k)
    assert lam.shape == (1,) and l.device.type == "cpu"
    assert rel(lam.item(), float(e[tag + "_lam"][0])) < EVAL_RTOL
    sgn = np.sign(np.dot(r.detach().numpy(), e[tag + "_r"]))
    assert np.abs(sgn * r.detach().numpy() - e[tag + "_r"]).max() < 1e-8
    assert np.abs(sgn * l.detach().numpy() - e[tag + "_l"]).max() < 1e-8 * np.abs(e[tag + "_l"]).max()
    assert abs(torch.dot(l, r).item() - 1.0) < 1e-12 and abs(torch.dot(r, r).item() - 1.0) < 1e-12
    loss = a * lam + l.matmul(M).matmul(r)
    gA, = torch.autograd.grad(loss, G)
    assert rel(loss.item(), float(e[tag + "_loss"])) < 1e-9
    assert np.abs(gA.numpy() - e[tag + "_gradA"]).max() < GRAD_RTOL * np.abs(e[tag + "_gradA"]).max()


def test_dominant_eig_reference_gradcheck(dsea, orc):
    """test_gradient.py:5-22 re-hosted: finite-difference gradcheck of DominantEig on a 25x25 transfer matrix."""
    from dominantsparseeigenad_b200.eig import DominantEig
    Gong = torch.from_numpy(orc.mps_transfer_matrix(5, 2, 7)).requires_grad_()
    torch.manual_seed(5)
    a = torch.randn(1, dtype=F64)
    Arandom = torch.randn(25, 25, dtype=F64)

    def func(A):
        eigval, l, r = DominantEig.apply(A, 25)
        return a * eigval + l.matmul(Arandom).matmul(r)

    assert torch.autograd.gradcheck(func, (Gong,), eps=1e-6, atol=1e-5, rtol=1e-4, nondet_tol=1e-7)


def test_dominant_sparse_eig_matches_dense(dsea, orc):
    """eig.py:64-152 / TFIM_vumps/general.py:57-75 shape: scipy LinearOperators + numpy adjoint callback."""
    import scipy.sparse.linalg as sla
    from dominantsparseeigenad_b200 import eig
    D, d = 6, 2
    rng = np.random.default_rng(11)
    A0 = rng.standard_normal((d, D, D))
    A = torch.from_numpy(A0).requires_grad_(True)
    Gong_t = torch.einsum("kij,kmn->imjn", A, A).reshape(D * D, D * D)
    lam_d, l_d, r_d = eig.DominantEig.apply(Gong_t, 30)
    M = torch.from_numpy(rng.standard_normal((D * D, D * D)))
    loss_d = 0.3 * lam_d + l_d.matmul(M).matmul(r_d)
    g_dense, = torch.autograd.grad(loss_d, A)

    fr = lambda v: np.einsum("kij,kmn,jn->im", A0, A0, v.reshape(D, D)).reshape(-1)
    fl = lambda v: np.einsum("kij,kmn,im->jn", A0, A0, v.reshape(D, D)).reshape(-1)
    Gong = sla.LinearOperator((D * D, D * D), matvec=fr)
    GongT = sla.LinearOperator((D * D, D * D), matvec=fl)

    def adj(grad_Gong):                                   # general.py:67-74
        gA = np.zeros((d, D, D))
        for u, v in grad_Gong:
            um, vm = u.reshape(D, D), v.reshape(D, D)
            gA = gA + np.einsum("im,jn,kmn->kij", um, vm, A0) + np.einsum("mi,nj,kmn->kij", um, vm, A0)
        return torch.from_numpy(gA)

    eig.setDominantSparseEig(Gong, GongT, adj)
    A2 = torch.from_numpy(A0).requires_grad_(True)
    lam_s, l_s, r_s = eig.DominantSparseEig.apply(A2, 30)
    loss_s = 0.3 * lam_s + l_s.matmul(M).matmul(r_s)
    g_sparse, = torch.autograd.grad(loss_s, A2)
    assert rel(lam_s.item(), lam_d.item()) < EVAL_RTOL and rel(loss_s.item(), loss_d.item()) < 1e-9
    assert (g_sparse - g_dense).abs().max().item() < GRAD_RTOL * g_dense.abs().max().item()


def test_dominant_eig_restarts_when_k_is_small(dsea):
    """k << n: the explicitly restarted Arnoldi must still converge to ARPACK-level accuracy."""
    from dominantsparseeigenad_b200.eig import DominantEig
    torch.manual_seed(9)
    n = 400
    A = torch.rand(n, n, dtype=F64) / n + torch.diag(torch.linspace(0.0, 1.0, n, dtype=F64))
    lam, l, r = DominantEig.apply(A, 20)
    w = torch.linalg.eigvals(A)
    ref = w[torch.argmax(w.abs())]
    assert abs(ref.imag.item()) < 1e-12 and rel(lam.item(), ref.real.item()) < EVAL_RTOL
    assert (A @ r - lam * r).norm().item() < 1e-9 and (A.T @ l - lam * l).norm().item() < 1e-9 * l.norm().item()


# ------------------------------------------------------------------------------------------------
# full-size invariants, edge cases, error behaviour
# ------------------------------------------------------------------------------------------------
def test_full_size_invariants_N24(dsea):
    """BASELINE's single-GPU size (2^24 amplitudes): properties that need no oracle.
    Ritz pair from k=48 vectors: |psi| = 1, Rayleigh quotient == theta, residual == beta_k |y_k| bound,
    basis orthonormal on sampled columns, operator symmetric."""
    N, k, g = 24, 48, 1.0
    m = dsea.TFIM(N)
    m.g = cuda([g])
    gen = torch.Generator(device="cuda").manual_seed(24)
    v = torch.randn(m.dim, dtype=F64, device="cuda", generator=gen)
    w = torch.randn(m.dim, dtype=F64, device="cuda", generator=gen)
    assert rel(torch.dot(w, m.H(v)).item(), torch.dot(m.H(w), v).item()) < 1e-11
    del v, w
    evals, psi, _, st = m.lanczos(m.g, k, 0)
    theta = evals[0].item()
    assert abs(torch.dot(psi, psi).item() - 1.0) < 1e-12
    Hpsi = m.H(psi)
    assert rel(torch.dot(psi, Hpsi).item(), theta) < 1e-12                       # Ritz value = Rayleigh quotient
    resid = (Hpsi - theta * psi).norm().item()
    assert resid <= st["beta"][: k - 1].max().item() * 1.0001                    # |r| = beta_k |y_k| <= max beta
    Q, ldq = st["Q"], st["ldq"]
    cols = [0, 1, k // 2, k - 2, k - 1]
    G = torch.stack([Q[j * ldq:j * ldq + m.dim] for j in cols])
    assert (G @ G.T - torch.eye(len(cols), dtype=F64, device="cuda")).abs().max().item() < 1e-12
    # upper bound property: theta_min(k) >= E0, and within 1e-3 already at k=48
    from oracle import dsea_oracle as orc
    E0 = orc.tfim_analytic(N, g)[0]
    assert theta >= E0 - 1e-9 and rel(theta, E0) < 1e-3


@pytest.mark.parametrize("k", [1, 2, 3])
def test_lanczos_tiny_k(dsea, k):
    torch.manual_seed(k)
    n = 50
    A = torch.randn(n, n, dtype=F64, device="cuda")
    A = A + A.T
    q0 = torch.randn(n, dtype=F64)
    dsea.runtime.set_start_vector_hook(lambda nn, kind: q0)
    try:
        lo, vlo, hi, vhi = dsea.Lanczos.symeigLanczos(A, k, device=torch.device("cuda"))
    finally:
        dsea.runtime.set_start_vector_hook(None)
    # reference semantics: eigenvalues of the k x k projection onto the Krylov space of q0
    Qs = [q0.cuda() / q0.norm()]
    for _ in range(k - 1):
        r = A @ Qs[-1]
        for q in Qs:
            r = r - torch.dot(q, r) * q
        Qs.append(r / r.norm())
    Qm = torch.stack(Qs, 1)
    w = torch.linalg.eigvalsh(Qm.T @ A @ Qm)
    assert rel(lo.item(), w[0].item()) < 1e-10 and rel(hi.item(), w[-1].item()) < 1e-10
    assert abs(vlo.norm().item() - 1) < 1e-12


def test_csr_operator_ragged_and_dense_rows(dsea):
    """CSR SpMV: odd dimension, empty rows, short rows (thread-per-row) and long rows (warp-per-row)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    for n, density in ((777, 0.004), (1001, 0.05)):
        M = sp.random(n, n, density=density, random_state=rng, format="csr")
        M = (M + M.T).tocsr()
        M[5, :] = 0
        M[:, 5] = 0
        M.eliminate_zeros()
        pot = torch.from_numpy(rng.standard_normal(n)).cuda()
        op = dsea.SparseMatrixOperator.from_scipy(M, pot)
        v = rng.standard_normal(n)
        want = M @ v + pot.cpu().numpy() * v
        got = op.H(torch.from_numpy(v).cuda()).cpu().numpy()
        assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
        # and through the solver: smallest eigenvalue of M + diag(pot)
        E0, psi = dsea.Lanczos.symeigLanczos(op.H, min(n, 400), device=torch.device("cuda"), extreme="min",
                                             sparse=True, dim=n)
        ref = np.linalg.eigvalsh(M.toarray() + np.diag(pot.cpu().numpy()))[0]
        assert rel(E0.item(), ref) < 1e-9


def test_error_behaviour(dsea):
    """Bad arguments raise (status + dsea_last_error), they do not crash or fall back."""
    from dominantsparseeigenad_b200 import _lib
    rt = dsea.runtime.context()
    with pytest.raises(_lib.DseaError):
        dsea.TFIM(1)                                         # too few spins
    with pytest.raises(_lib.DseaError):
        dsea.TFIM(64)
    A = torch.eye(8, dtype=F64, device="cuda")
    with pytest.raises(_lib.DseaError):
        dsea.Lanczos.symeigLanczos(A, 5000)                  # k beyond the supported maximum
    with pytest.raises(ValueError):
        dsea.Lanczos.symeigLanczos(A, 4, extreme="middle")
    with pytest.raises(TypeError):
        dsea.Lanczos.symeigLanczos(lambda v: v, 4)           # callable without sparse=True
    with pytest.raises(_lib.DseaError):
        rt.set_option("no_such_option", 1)
    m = dsea.TFIM(6)
    with pytest.raises(ValueError):
        m.H(torch.zeros(64, dtype=F64, device="cuda"))       # g not set


# ------------------------------------------------------------------------------------------------
# round 2: production sweep kernels element-wise, config 3 at full size, non-convergence, CPU callers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [20, 22])
def test_tfim_production_sweeps_elementwise_vs_oracle(dsea, orc, N):
    """H, dH/dg and the adjoint at N = 20 / 22 with the DEFAULT tile plan, element by element against the
    oracle's table-driven operator (TFIM.py:91-101 restated): at these sizes the second sweep takes the
    persistent strided kernel and the first one the TMA-staged kernel, which the N <= 18 plan sweep
    never reaches.  (The oracle's (2^N, N) int64 table is 168 MB / 738 MB here.)"""
    rt = dsea.runtime.context()
    rt.set_option("tfim_tile_bits", 13)
    rt.set_option("tfim_run_bits", 0)
    g = 0.75 + 0.01 * N
    rng = np.random.default_rng(100 + N)
    v, w = rng.standard_normal(1 << N), rng.standard_normal(1 << N)
    o = orc.TFIMOracle(N, g)
    m = dsea.TFIM(N)
    m.g = cuda([g])
    vd, wd = cuda(v), cuda(w)
    want = o.H(torch.from_numpy(v)).numpy()
    got = m.H(vd).cpu().numpy()
    # every output is a sum of N + 1 terms of magnitude <= (N + g) max|v|: summation order differs, nothing else
    assert np.abs(got - want).max() <= 4e-16 * (N + 1) * (N + g) * np.abs(v).max()
    wantp = o.pHpg(torch.from_numpy(v)).numpy()
    gotp = m.pHpg(vd).cpu().numpy()
    assert np.abs(gotp - wantp).max() <= 4e-16 * N * N * np.abs(v).max()
    want_adj = o.Hadjoint_to_gadjoint(torch.from_numpy(w), torch.from_numpy(v)).item()
    assert rel(m.Hadjoint_to_gadjoint(wd, vd).item(), want_adj) < 1e-11
    # shifted operator with the dot epilogue, as the CG loop calls it
    u = m.matvec_raw(m.g, vd, cuda([-3.25]))
    assert np.abs(u.cpu().numpy() - (want + 3.25 * v)).max() <= 4e-16 * (N + 2) * (N + g + 3.25) * np.abs(v).max()


@pytest.mark.parametrize("g", [1.0, 1.5])
def test_config3_chiF_and_d2E0_N24_k200(dsea, g):
    """BASELINE config 3 at full size on one GPU: N = 24, k = 200, fp64 — E0, dE0/dg, d2E0/dg2 and the fidelity
    susceptibility (three nested CG solves, CG.py:128-138) against the closed forms of BASELINE.md section 3
    (chi_F = 17.25 / 0.534180712, d2E0/N = -1.174367973 / -0.183237299)."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    N, k = 24, 200
    ex = tfim_exact(N, g)
    E0, dE0, d2E0, psi0 = _tfim_E0_family(dsea, N, g, k)
    assert rel(E0, ex.E0) < EVAL_RTOL
    assert rel(dE0, ex.dE0) < GRAD_RTOL
    assert rel(d2E0, ex.d2E0) < GRAD_RTOL
    assert abs(dsea.dot(psi0, psi0).item() - 1.0) < 1e-12
    del psi0
    torch.cuda.empty_cache()
    chif = _tfim_chif(dsea, N, g, k)
    assert rel(chif, ex.chiF) < GRAD_RTOL
    if g == 1.0:
        assert rel(chif, 17.25) < GRAD_RTOL and rel(d2E0 / N, -1.174367973012996) < GRAD_RTOL
    torch.cuda.empty_cache()


def test_cg_reports_non_convergence(dsea):
    """The reference returns silently after n iterations (CG.py:32).  The C ABI returns DSEA_ERR_NOCONV and the
    Python layer turns it into a ConvergenceWarning while still handing back the last iterate."""
    from dominantsparseeigenad_b200 import _lib
    n = 300
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n))
    A = A @ A.T / n + np.eye(n)                           # SPD, condition number ~5: converges in a few dozen steps
    b = rng.standard_normal(n)
    op = dsea.DenseOperator(cuda(A))
    with pytest.warns(_lib.ConvergenceWarning):
        x = op.cg(None, None, cuda(b), cuda(np.zeros(n)), maxit=3)
    assert dsea.runtime.stats["cg_iters"][-1] == 3 and torch.isfinite(x).all()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error", _lib.ConvergenceWarning)
        x = op.cg(None, None, cuda(b), cuda(np.zeros(n)))     # default cap n: converges, no warning
    assert np.abs(A @ x.cpu().numpy() - b).max() < 1e-6


def test_second_derivative_from_cpu_tensors(dsea, orc):
    """CPU tensors in => CPU tensors out in every order of differentiation (the reference's documented contract,
    CG.py docstring): d2E0/dg2 through DominantSymeig on a CPU matrix, against torch's own AD through eigh."""
    N, g0 = 6, 1.1
    o = orc.TFIMOracle(N, 1.0)
    D = torch.diag(o.diag_elements.clone())
    X = torch.zeros(o.dim, o.dim, dtype=F64)
    cols = torch.arange(o.dim)
    for i in range(N):
        X[o.flips_basis[:, i], cols] -= 1.0

    def second(fn):
        g = torch.tensor([g0], dtype=F64, requires_grad=True)
        E0 = fn(D + g * X)
        dE0, = torch.autograd.grad(E0, g, create_graph=True)
        d2E0, = torch.autograd.grad(dE0, g)
        assert E0.device.type == "cpu" and dE0.device.type == "cpu" and d2E0.device.type == "cpu"
        return E0.item(), dE0.item(), d2E0.item()

    ours = second(lambda H: dsea.symeig.DominantSymeig.apply(H, 64)[0])
    # eigh's AD divides by eigenvalue gaps; the even/odd sectors of H never mix, so use the analytic values too
    aE0, adE0, ad2E0, _ = orc.tfim_analytic(N, g0)
    assert rel(ours[0], aE0) < EVAL_RTOL and rel(ours[1], adE0) < GRAD_RTOL and rel(ours[2], ad2E0) < GRAD_RTOL


def test_selfcheck_single_gpu(dsea):
    """The identities bench.py asserts at P > 1 (character vectors per spin bit, <1,H1>, symmetry, small solve)
    hold on one GPU as well — every bit class of the sweep plan (register, shared memory, strided, direct)."""
    from dominantsparseeigenad_b200 import selfcheck
    res = selfcheck.run(model=dsea.TFIM(23), small_N=14)
    assert res["ok"], res


def test_start_vector_stream_restarts_with_the_seed(dsea):
    torch.manual_seed(4321)
    a = dsea.runtime.start_vector(1000, "lanczos").clone()
    b = dsea.runtime.start_vector(1000, "lanczos").clone()
    torch.manual_seed(99)
    c = dsea.runtime.start_vector(1000, "lanczos").clone()
    torch.manual_seed(4321)
    a2 = dsea.runtime.start_vector(1000, "lanczos").clone()
    b2 = dsea.runtime.start_vector(1000, "lanczos").clone()
    assert torch.equal(a, a2) and torch.equal(b, b2) and not torch.equal(a, b) and not torch.equal(a, c)


# ------------------------------------------------------------------------------------------------
# opt-in fp32 shadow basis (capacity for N = 30, k = 200; half the reorth traffic)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,k,g", [(16, 200, 1.0), (16, 200, 1.5), (20, 100, 1.25), (13, 150, 1.0)])
def test_fp32_shadow_basis_keeps_the_fp64_tolerances(dsea, N, k, g):
    """Lanczos vectors stored rounded to fp32 + one fp64 Jacobi-Davidson polish: E0 to 1e-10, overlap with the
    fp64-basis eigenvector >= 1 - 1e-8, dE0 / d2E0 / chi_F to 1e-6 — the north-star tolerances — and an eigen-residual
    no worse than the fp64 path's."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    ex = tfim_exact(N, g)
    E64, dE64, d2E64, psi64 = _tfim_E0_family(dsea, N, g, k)
    dsea.runtime.set_basis_precision("fp32")
    try:
        E32, dE32, d2E32, psi32 = _tfim_E0_family(dsea, N, g, k)
        chif = _tfim_chif(dsea, N, g, k)
        m = dsea.TFIM(N)
        m.g = cuda([g])
        resid32 = (m.H(psi32) - E32 * psi32).norm().item()
    finally:
        dsea.runtime.set_basis_precision("fp64")
    assert rel(E32, ex.E0) < EVAL_RTOL and rel(E32, E64) < EVAL_RTOL
    assert 1 - abs(torch.dot(psi32, psi64).item()) < OVERLAP_TOL
    assert abs(psi32.norm().item() - 1.0) < 1e-12
    assert rel(dE32, ex.dE0) < GRAD_RTOL and rel(d2E32, ex.d2E0) < GRAD_RTOL and rel(chif, ex.chiF) < GRAD_RTOL
    assert resid32 < 1e-8, resid32


def test_fp32_shadow_basis_full_size_N24_k200(dsea):
    """Config 3's size with the compressed basis (13.4 GB instead of 26.8 GB): E0, dE0/dg against the closed forms."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    N, k, g = 24, 200, 1.0
    ex = tfim_exact(N, g)
    dsea.runtime.set_basis_precision("fp32")
    try:
        m = dsea.TFIM(N)
        m.g = torch.tensor([g], dtype=F64, device="cuda", requires_grad=True)
        dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
        E0, psi0 = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, torch.device("cuda"))
        dE0, = torch.autograd.grad(E0, m.g)
        resid = (m.H(psi0.detach()) - E0.detach() * psi0.detach()).norm().item()
    finally:
        dsea.runtime.set_basis_precision("fp64")
    assert rel(E0.item(), ex.E0) < EVAL_RTOL and rel(dE0.item(), ex.dE0) < GRAD_RTOL and resid < 1e-8
    torch.cuda.empty_cache()


def test_fp32_shadow_basis_is_min_only_and_tfim_only(dsea):
    from dominantsparseeigenad_b200 import _lib
    dsea.runtime.set_basis_precision("fp32")
    try:
        m = dsea.TFIM(10)
        m.g = cuda([1.0])
        with pytest.raises(_lib.DseaError):
            dsea.Lanczos.symeigLanczos(m.H, 40, device=torch.device("cuda"), extreme="both", sparse=True, dim=m.dim)
        A = torch.randn(64, 64, dtype=F64, device="cuda")
        A = A + A.T
        lo, v = dsea.Lanczos.symeigLanczos(A, 64, extreme="min")       # dense operators keep the fp64 basis
        assert rel(lo.item(), torch.linalg.eigvalsh(A)[0].item()) < EVAL_RTOL
    finally:
        dsea.runtime.set_basis_precision("fp64")


# ------------------------------------------------------------------------------------------------
# callers either side of the path: the example drivers are exercised, not just shipped
# ------------------------------------------------------------------------------------------------
def _load_example(name):
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("dsea_example_" + name, os.path.join(ROOT, "examples", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_vumps_consumer_dense_and_matrix_free_agree_and_descend(dsea):
    """examples/tfim_vumps.py (the reference's TFIM_vumps/general.py consumer of DominantEig / DominantSparseEig):
    both forward variants give the same energy and gradient, the gradient matches finite differences, and a few
    L-BFGS steps approach the exact infinite-chain energy from above."""
    ex = _load_example("tfim_vumps")
    g, D = 1.0, 4
    model = ex.UniformMPS(D, g, k=D * D, seed=3)
    e_dense = model.energy_dense()
    grad_dense, = torch.autograd.grad(e_dense, model.A)
    e_free = model.energy_matrix_free()
    grad_free, = torch.autograd.grad(e_free, model.A)
    assert rel(e_free.item(), e_dense.item()) < 1e-10
    assert (grad_free - grad_dense).abs().max().item() < 1e-7 * max(1.0, grad_dense.abs().max().item())
    # finite differences along a random direction
    gen = torch.Generator().manual_seed(5)
    dirn = torch.randn(2, D, D, dtype=F64, generator=gen).cuda()
    eps = 1e-5
    with torch.no_grad():
        A0 = model.A.detach().clone()
        model.A.copy_(A0 + eps * dirn)
        ep = model.energy_dense().item()
        model.A.copy_(A0 - eps * dirn)
        em = model.energy_dense().item()
        model.A.copy_(A0)
    fd = (ep - em) / (2 * eps)
    assert rel((grad_dense * dirn).sum().item(), fd) < 1e-6
    hist = ex.optimise(model, steps=12, sparse=False, verbose=False)
    e0 = ex.exact_energy_per_site(g)
    assert hist[-1] < hist[0] and hist[-1] >= e0 - 1e-9 and hist[-1] - e0 < 5e-3, (hist[0], hist[-1], e0)
    assert abs(e0 + 4.0 / np.pi) < 1e-8                       # g = 1: e0 = -4 / pi


def test_example_drivers_run(dsea, capsys):
    """examples/tfim_E0.py, tfim_chiF.py (E0.py:94-113, chiF.py:65-82 shaped sweep drivers) and schrodinger1D.py run
    end to end at a small size and print values that agree with the closed forms."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    import os
    import tempfile
    e0 = _load_example("tfim_E0")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "E0_N_10.npz")
        rows = e0.sweep(N=10, k=120, gs=[1.0, 1.25], save=path)
        saved = np.load(path)
        assert sorted(saved.files) == ["E0s", "d2E0s", "dE0s", "gs"]          # the keys of examples/TFIM/datas/E0_N_*.npz
        for (g, E0, dE0, d2E0), sE0 in zip(rows, saved["E0s"]):
            ex = tfim_exact(10, g)
            assert rel(E0 * 10, ex.E0) < EVAL_RTOL and rel(dE0 * 10, ex.dE0) < GRAD_RTOL and rel(d2E0 * 10, ex.d2E0) < GRAD_RTOL
            assert sE0 == E0
        chi = _load_example("tfim_chiF")
        cpath = os.path.join(tmp, "chiF_N_10.npz")
        for g, c in chi.sweep(N=10, k=120, gs=[1.25], save=cpath):
            assert rel(c, tfim_exact(10, g).chiF) < GRAD_RTOL
        assert sorted(np.load(cpath).files) == ["chiFs", "gs"]
    sch = _load_example("schrodinger1D")
    hist = sch.fit("csr", N=300, k=300, steps=3, verbose=False)           # config 1 as shipped, native CSR operator
    assert abs(hist[0] - 0.099454537767) < 1e-9 and hist[-1] < 0.02       # 0.0995 -> 0.0168 -> 0.0097


@pytest.mark.parametrize("N,k", [(6, 64), (9, 200), (11, 130)])
def test_fp32_shadow_basis_ragged_tiles(dsea, N, k):
    """n_loc below / not a multiple of the 2048-row fp32 reorth tile: the scalar tail paths of both passes and of the
    rounding store, with k up to the full dimension (breakdown handling)."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    g = 1.3
    ex = tfim_exact(N, g)
    dsea.runtime.set_basis_precision("fp32")
    try:
        E0, dE0, d2E0, psi = _tfim_E0_family(dsea, N, g, min(k, 1 << N))
    finally:
        dsea.runtime.set_basis_precision("fp64")
    assert rel(E0, ex.E0) < EVAL_RTOL and rel(dE0, ex.dE0) < GRAD_RTOL and rel(d2E0, ex.d2E0) < GRAD_RTOL
    assert abs(psi.norm().item() - 1.0) < 1e-12


def test_zero_start_cg_matches_random_start(dsea):
    """Opt-in x0 = 0 for the subspace solves: same dE0 / d2E0 / chi_F (the solution on psi-perp is unique), and the
    E0-only backward (zero right-hand side) needs no iteration at all."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    N, k, g = 16, 160, 1.25
    ex = tfim_exact(N, g)
    dsea.runtime.cg_start = "zero"
    try:
        dsea.runtime.stats["cg_iters"].clear()
        E0, dE0, d2E0, _ = _tfim_E0_family(dsea, N, g, k)
        iters = list(dsea.runtime.stats["cg_iters"])
        chif = _tfim_chif(dsea, N, g, k)
    finally:
        dsea.runtime.cg_start = "random"
    assert iters[0] == 0                                     # first backward: b = 0
    assert rel(E0, ex.E0) < EVAL_RTOL and rel(dE0, ex.dE0) < GRAD_RTOL and rel(d2E0, ex.d2E0) < GRAD_RTOL
    assert rel(chif, ex.chiF) < GRAD_RTOL


def test_even_parity_start_vectors_cure_the_g_below_1_scatter(dsea):
    """g = 0.5: the odd-parity partner of the ground state is only ~g^N above it, and with generic start vectors the
    reference's own d2E0 / chi_F scatter by 1e-4 run to run.  With the opt-in even-sector start vectors the
    second-order quantities meet the 1e-6 tolerance there as well, for different seeds."""
    from dominantsparseeigenad_b200.analytic import tfim_exact
    N, k, g = 16, 200, 0.5
    ex = tfim_exact(N, g)
    dsea.runtime.parity_sector = "even"
    try:
        for seed in (1, 2):
            torch.manual_seed(seed)
            E0, dE0, d2E0, psi = _tfim_E0_family(dsea, N, g, k)
            chif = _tfim_chif(dsea, N, g, k)
            assert rel(E0, ex.E0) < EVAL_RTOL and rel(dE0, ex.dE0) < GRAD_RTOL
            assert rel(d2E0, ex.d2E0) < GRAD_RTOL and rel(chif, ex.chiF) < GRAD_RTOL, (seed, d2E0, ex.d2E0, chif, ex.chiF)
            assert (psi - psi.flip(0)).abs().max().item() < 1e-9          # the eigenvector is even
    finally:
        dsea.runtime.parity_sector = "none"
