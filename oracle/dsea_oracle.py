"""CPU restatement of the reference's dominant-eigenpair path (TEST INFRASTRUCTURE).

Nothing here is on the product path; see oracle/__init__.py.  Citations are
`file:line` under the reference checkout (/root/reference).  Integer/bit work is
numpy; fp64 vector work is torch-CPU so that (a) second derivatives flow through
torch.autograd exactly the way they do in the reference and (b) the CPU baseline
timed by bench.py uses the same library kernels, thread pool and memory traffic
as the reference's own PyTorch CPU path.

Randomness: the reference draws unseeded `torch.randn` start vectors
(Lanczos.py:52,59; CG.py:58,121).  Here every draw goes through a `draw(n)`
callable so tests can feed both sides the same vectors.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch

__all__ = [
    "SeededDraws", "ListDraws", "tfim_flip_table", "tfim_diagonal", "tfim_diagonal_closed_form",
    "TFIMOracle", "lanczos_basis", "extreme_eigpair", "cg_solve", "make_sparse_primitives",
    "dense_dominant_symeig", "tfim_analytic", "tfim_energy_derivatives", "tfim_fidelity_susceptibility",
    "Schrodinger1DOracle", "MatvecCounter", "dominant_eig_triple", "dominant_eig_backward", "mps_transfer_matrix",
]

F64 = torch.float64
_TORCH_RANDN = torch.randn      # captured so gen_golden's replay patch of torch.randn cannot recurse


# ----------------------------------------------------------------------------------------------
# start-vector sources
# ----------------------------------------------------------------------------------------------
class SeededDraws:
    """Deterministic stand-in for the reference's unseeded torch.randn draws."""

    def __init__(self, seed: int):
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(int(seed))
        self.count = 0

    def __call__(self, n: int) -> torch.Tensor:
        self.count += 1
        return _TORCH_RANDN(n, dtype=F64, generator=self.gen)


class ListDraws:
    """Replays a fixed list of vectors (used to inject identical q0 / x0 everywhere)."""

    def __init__(self, vectors: Sequence[torch.Tensor]):
        self.vectors = [torch.as_tensor(v, dtype=F64).clone() for v in vectors]
        self.count = 0

    def __call__(self, n: int) -> torch.Tensor:
        v = self.vectors[self.count]
        self.count += 1
        assert v.shape == (n,), (v.shape, n)
        return v.clone()


class MatvecCounter:
    """Wraps an operator callable and counts applications (work accounting)."""

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor]):
        self.fn = fn
        self.calls = 0

    def __call__(self, v):
        self.calls += 1
        return self.fn(v)


# ----------------------------------------------------------------------------------------------
# TFIM integer tables (bit-exact contract)
# ----------------------------------------------------------------------------------------------
def tfim_flip_table(N: int) -> np.ndarray:
    """flips[s, i] = s XOR (1 << i), int64, shape (2^N, N).   Restates TFIM.py:48-51.

    The reference builds the masks through a float32 tensor before `.long()`; powers of two
    below 2^31 are exact in float32, so plain integer shifts give the identical table.
    """
    s = np.arange(1 << N, dtype=np.int64)[:, None]
    masks = (np.int64(1) << np.arange(N, dtype=np.int64))[None, :]
    return s ^ masks


def tfim_diagonal(N: int) -> np.ndarray:
    """diag[s] = -sum_j z_j z_{j+1 mod N},  z_j = 1 - 2*bit_{N-1-j}(s).   Restates TFIM.py:39-46.

    Built column by column (O(2^N) memory) rather than as the reference's (2^N, N) table; the
    value is an exact small integer so summation order is irrelevant.  Returned as float64.
    """
    n = 1 << N
    s = np.arange(n, dtype=np.int64)
    acc = np.zeros(n, dtype=np.int64)
    for j in range(N):
        bit_j = (s >> (N - 1 - j)) & 1                      # spin j <-> bit N-1-j   (TFIM.py:41)
        bit_n = (s >> (N - 1 - ((j + 1) % N))) & 1          # periodic neighbour     (TFIM.py:43)
        acc += (1 - 2 * bit_j) * (1 - 2 * bit_n)            # TFIM.py:42,44
    return (-acc).astype(np.float64)


def tfim_diagonal_closed_form(N: int, s: np.ndarray) -> np.ndarray:
    """-(N - 2*popcount(s ^ rotl_N(s, 1))): the bit-arithmetic form the CUDA kernels use.

    z_j z_{j+1} = +1 when the two bits agree and -1 otherwise, so the bond sum is
    N - 2*(number of differing cyclic neighbours) = N - 2*popcount(s ^ rotl_N(s)).
    Proven equal to tfim_diagonal() by tests/test_oracle.py for every s at N<=16.
    """
    s = np.asarray(s, dtype=np.uint64)
    full = np.uint64((1 << N) - 1)
    rot = ((s << np.uint64(1)) | (s >> np.uint64(N - 1))) & full
    x = s ^ rot
    pop = np.zeros(x.shape, dtype=np.int64)
    for b in range(N):
        pop += ((x >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
    if N == 1:      # single site, bond with itself (never used; keeps the formula total)
        return -np.ones(x.shape)
    return (-(N - 2 * pop)).astype(np.float64)


class TFIMOracle:
    """H = -sum_i (g sx_i + sz_i sz_{i+1}), periodic.  Restates TFIM.py:5-16, 58-65, 91-101.

    Holds the (2^N, N) int64 flip table exactly as the reference does, so its matvec moves the
    same bytes as the reference's (this is what the CPU baseline is supposed to measure).
    """

    def __init__(self, N: int, g: float | torch.Tensor = 1.0):
        self.N = int(N)
        self.dim = 1 << self.N
        self.diag_elements = torch.from_numpy(tfim_diagonal(self.N))
        self.flips_basis = torch.from_numpy(tfim_flip_table(self.N))
        self.g = g if isinstance(g, torch.Tensor) else torch.tensor([float(g)], dtype=F64)

    def H(self, v: torch.Tensor) -> torch.Tensor:                       # TFIM.py:91-98
        return v * self.diag_elements - self.g * v[self.flips_basis].sum(dim=1)

    def pHpg(self, v: torch.Tensor) -> torch.Tensor:                    # TFIM.py:58-65
        return -v[self.flips_basis].sum(dim=1)

    def Hadjoint_to_gadjoint(self, v1, v2) -> torch.Tensor:             # TFIM.py:100-101
        return self.pHpg(v2).matmul(v1)[None]

    def dense(self) -> torch.Tensor:
        """Dense H (no noise term; TFIM.py:67-89 adds 1e-12 noise which we omit)."""
        n = self.dim
        Hm = torch.diag(self.diag_elements.clone())
        cols = torch.arange(n)
        for i in range(self.N):
            Hm[self.flips_basis[:, i], cols] -= float(self.g)
        return Hm


# ----------------------------------------------------------------------------------------------
# Lanczos with full re-orthogonalisation
# ----------------------------------------------------------------------------------------------
def lanczos_basis(Amap, n: int, k: int, draw) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """k-step Lanczos, one classical Gram-Schmidt sweep per step.  Restates Lanczos.py:49-77.

    Returns (Q (n,k) row-major like the reference, alphas (k,), betas (k-1,)).
    """
    Q = torch.zeros((n, k), dtype=F64)
    alphas = torch.zeros(k, dtype=F64)
    betas = torch.zeros(max(k - 1, 0), dtype=F64)
    q = draw(n)
    q = q / torch.norm(q)                                    # Lanczos.py:52-53
    u = Amap(q)                                              # :54
    alpha = torch.dot(q, u)                                  # :55
    Q[:, 0] = q
    alphas[0] = alpha
    beta = 0.0
    q_prev = draw(n)                                         # :59 (drawn, then multiplied by beta=0)
    for i in range(1, k):
        r = u - alpha * q - beta * q_prev                    # :61
        basis = Q[:, :i]
        r = r - basis.matmul(basis.T.matmul(r))              # :66 full reorth, single CGS sweep
        q_prev = q
        beta = torch.norm(r)                                 # :69
        q = r / beta                                         # :70
        u = Amap(q)                                          # :71
        alpha = torch.dot(q, u)                              # :72
        alphas[i] = alpha
        betas[i - 1] = beta
        Q[:, i] = q
    return Q, alphas, betas


def extreme_eigpair(Amap, n: int, k: int, draw, which: str = "min"):
    """Ritz pair(s) from the Lanczos tridiagonal.  Restates Lanczos.py:97-105.

    `torch.symeig` (Lanczos.py:98) no longer exists; `torch.linalg.eigh` on the same dense T is
    the documented replacement (ascending eigenvalues, orthonormal columns).
    """
    Q, a, b = lanczos_basis(Amap, n, k, draw)
    T = torch.diag(a) + torch.diag(b, 1) + torch.diag(b, -1)           # Lanczos.py:76
    w, Y = torch.linalg.eigh(T)
    X = Q.matmul(Y)                                                     # :99 (all k Ritz vectors)
    if which == "min":
        return w[0], X[:, 0]
    if which == "max":
        return w[-1], X[:, -1]
    return w[0], X[:, 0], w[-1], X[:, -1]


# ----------------------------------------------------------------------------------------------
# CG
# ----------------------------------------------------------------------------------------------
def cg_solve(Amap, b: torch.Tensor, x0: torch.Tensor, eps: float = 1e-7, info: Optional[dict] = None):
    """Plain CG with the reference's stopping rule |r| < 1e-7 absolute.  Restates CG.py:22-41.

    Like the reference it applies the operator twice per iteration (CG.py:34 and :40); that is
    the cost the CPU baseline reports.  `info["iters"]` receives the iteration count.
    """
    n = b.shape[0]
    x = x0
    r = b - Amap(x)                                          # CG.py:27
    iters = 0
    if torch.norm(r).item() >= eps:                          # :28
        d = r
        alpha = torch.dot(r, r) / torch.dot(Amap(d), d)      # :31
        for _ in range(n):                                   # :32
            iters += 1
            x = x + alpha * d                                # :33
            r_next = r - alpha * Amap(d)                     # :34
            if torch.norm(r_next).item() < eps:              # :35
                break
            beta = torch.dot(r_next, r_next) / torch.dot(r, r)   # :37
            r = r_next
            d = r + beta * d                                 # :39
            alpha = torch.dot(r, r) / torch.dot(Amap(d), d)  # :40
    if info is not None:
        info["iters"] = iters
    return x


# ----------------------------------------------------------------------------------------------
# Differentiable primitives (matrix-free)
# ----------------------------------------------------------------------------------------------
def make_sparse_primitives(A, Aadjoint_to_padjoint, draw, stats: Optional[dict] = None):
    """Builds the pair (DominantSparseSymeig, CGSubspaceSparse) for an operator closure.

    Restates symeig.py:33-88 and CG.py:73-140.  Unlike the reference these are returned rather
    than bound to module globals, and the start vectors come from `draw`.
    """
    stats = stats if stats is not None else {}
    stats.setdefault("cg_iters", [])

    class CGSubspaceSparse(torch.autograd.Function):
        @staticmethod
        def forward(ctx, p, E0, b, psi):                                  # CG.py:118-126
            shifted = lambda v: A(v) - E0 * v
            x0 = draw(b.shape[0])
            x0 = x0 - torch.dot(psi, x0) * psi                            # :122
            info = {}
            x = cg_solve(shifted, b, x0, info=info)
            stats["cg_iters"].append(info["iters"])
            ctx.p = p
            ctx.save_for_backward(E0, psi, x)
            return x

        @staticmethod
        def backward(ctx, grad_x):                                        # CG.py:128-138
            E0, psi, x = ctx.saved_tensors
            rhs = grad_x - torch.dot(psi, grad_x) * psi                   # :132
            grad_b = CGSubspaceSparse.apply(ctx.p, E0, rhs, psi)          # :133
            v1, v2 = -grad_b, x
            grad_psi = -x * torch.dot(psi, grad_x)                        # :135
            grad_E0 = -torch.dot(v1, v2)                                  # :136
            grad_p = Aadjoint_to_padjoint(v1, v2)                         # :137
            return grad_p, grad_E0, grad_b, grad_psi

    class DominantSparseSymeig(torch.autograd.Function):
        @staticmethod
        def forward(ctx, p, k, dim):                                      # symeig.py:71-75
            E0, psi = extreme_eigpair(A, dim, k, draw, "min")
            ctx.save_for_backward(p, E0, psi)
            return E0, psi

        @staticmethod
        def backward(ctx, grad_E0, grad_psi):                             # symeig.py:77-86
            p, E0, psi = ctx.saved_tensors
            rhs = grad_psi - torch.dot(psi, grad_psi) * psi               # :80
            lam = CGSubspaceSparse.apply(p, E0, rhs, psi)                 # :81
            v1, v2 = grad_E0 * psi - lam, psi                             # :82-83
            return Aadjoint_to_padjoint(v1, v2), None, None

    return DominantSparseSymeig, CGSubspaceSparse


# ----------------------------------------------------------------------------------------------
# Differentiable primitives (dense)
# ----------------------------------------------------------------------------------------------
def dense_dominant_symeig(draw):
    """Builds (DominantSymeig, CGSubspace) for dense tensors.  Restates symeig.py:4-31, CG.py:43-71."""

    class CGSubspace(torch.autograd.Function):
        @staticmethod
        def forward(ctx, M, b, psi):                                      # CG.py:56-62
            x0 = draw(b.shape[0]).to(b.dtype)
            x0 = x0 - torch.dot(psi, x0) * psi
            x = cg_solve(lambda v: M.matmul(v), b, x0)
            ctx.save_for_backward(M, psi, x)
            return x

        @staticmethod
        def backward(ctx, grad_x):                                        # CG.py:64-71
            M, psi, x = ctx.saved_tensors
            rhs = grad_x - torch.dot(psi, grad_x) * psi
            grad_b = CGSubspace.apply(M, rhs, psi)
            grad_M = -grad_b[:, None] * x                                 # :69
            grad_psi = -x * torch.dot(psi, grad_x)                        # :70
            return grad_M, grad_b, grad_psi

    class DominantSymeig(torch.autograd.Function):
        @staticmethod
        def forward(ctx, M, k):                                           # symeig.py:15-19
            E0, psi = extreme_eigpair(lambda v: M.matmul(v), M.shape[0], k, draw, "min")
            ctx.save_for_backward(M, E0, psi)
            return E0, psi

        @staticmethod
        def backward(ctx, grad_E0, grad_psi):                             # symeig.py:21-31
            M, E0, psi = ctx.saved_tensors
            shifted = M - E0 * torch.eye(M.shape[0], dtype=M.dtype)       # :25
            rhs = grad_psi - torch.dot(psi, grad_psi) * psi               # :27
            lam = CGSubspace.apply(shifted, rhs, psi)                     # :28
            grad_M = (grad_E0 * psi - lam)[:, None] * psi                 # :29
            return grad_M, None

    return DominantSymeig, CGSubspace


# ----------------------------------------------------------------------------------------------
# Analytic TFIM results (exact at finite N, Neveu-Schwarz momenta)
# ----------------------------------------------------------------------------------------------
def tfim_analytic(N: int, g: float):
    """Total (not per-site) E0, dE0/dg, d2E0/dg2 and chi_F.  E0 part restates E0.py:15-20.

    Momenta are the Neveu-Schwarz set k = (2m+1) pi / N (even fermion parity, where the ground state
    lives).  For even N this IS the reference's grid (m - (N-1)/2) 2 pi / N; for odd N the reference's
    linspace lands on the Ramond set (multiples of 2 pi / N) and is off by O(1/N^2) — verified against
    the CUDA solver at N=25 (NS agrees to 1e-16, the reference grid differs by 2e-3).
    chi_F = 1/4 sum_{k>0} sin^2 k / (1 + g^2 - 2 g cos k)^2 is not in the reference.
    """
    m = np.arange(N, dtype=np.float64)
    ks = (2.0 * m + 1.0) * math.pi / N
    ks = np.where(ks > math.pi + 1e-12, ks - 2.0 * math.pi, ks)
    eps = 2.0 * np.sqrt(g * g - 2.0 * g * np.cos(ks) + 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        d_eps = np.where(eps > 0, 4.0 * (g - np.cos(ks)) / eps, 0.0)
        d2_eps = np.where(eps > 0, 16.0 * np.sin(ks) ** 2 / eps ** 3, 0.0)
    pos = (ks > 0) & (ks < math.pi - 1e-12)
    chif = 0.25 * np.sum(np.sin(ks[pos]) ** 2 / (1.0 + g * g - 2.0 * g * np.cos(ks[pos])) ** 2)
    return -0.5 * eps.sum(), -0.5 * d_eps.sum(), -0.5 * d2_eps.sum(), float(chif)


# ----------------------------------------------------------------------------------------------
# Drivers shaped like the reference's example harnesses
# ----------------------------------------------------------------------------------------------
def tfim_energy_derivatives(model: TFIMOracle, k: int, draw, stats: Optional[dict] = None):
    """(E0, dE0/dg, d2E0/dg2) totals via the sparse primitive.  Restates E0.py:53-67."""
    Dom, _ = make_sparse_primitives(model.H, model.Hadjoint_to_gadjoint, draw, stats)
    E0, psi0 = Dom.apply(model.g, k, model.dim)
    dE0, = torch.autograd.grad(E0, model.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, model.g)
    return E0.item(), dE0.item(), d2E0.item(), psi0.detach()


def tfim_fidelity_susceptibility(model: TFIMOracle, k: int, draw, stats: Optional[dict] = None):
    """chi_F = -d2/dg2 log <psi0(g_fixed)|psi0(g)>.  Restates chiF.py:40-53."""
    Dom, _ = make_sparse_primitives(model.H, model.Hadjoint_to_gadjoint, draw, stats)
    E0, psi0 = Dom.apply(model.g, k, model.dim)
    logF = torch.log(psi0.detach().matmul(psi0))
    dlogF, = torch.autograd.grad(logF, model.g, create_graph=True)
    d2logF, = torch.autograd.grad(dlogF, model.g)
    return E0.item(), psi0.detach(), -d2logF.item()


class Schrodinger1DOracle:
    """1-D finite-difference Hamiltonian with a trainable potential.  Restates schrodinger1D.py:6-73,91-94."""

    def __init__(self, N: int = 300, xmin: float = -1.0, xmax: float = 1.0, potential=None):
        self.N = N
        x = np.linspace(xmin, xmax, num=N, endpoint=False)
        self.xmesh = torch.from_numpy(x)
        self.h = (xmax - xmin) / N
        self.potential = (0.5 * self.xmesh ** 2 if potential is None
                          else torch.as_tensor(potential, dtype=F64).clone()).requires_grad_(True)
        t = np.zeros(N)
        idx = np.abs(x) < 0.5
        t[idx] = 1.0 - np.abs(x[idx])
        self.target = torch.from_numpy(t / np.linalg.norm(t))               # :91-94

    def kinetic_dense(self) -> torch.Tensor:                                # :12-15
        N = self.N
        return -0.5 / self.h ** 2 * (torch.diag(-2 * torch.ones(N, dtype=F64))
                                     + torch.diag(torch.ones(N - 1, dtype=F64), 1)
                                     + torch.diag(torch.ones(N - 1, dtype=F64), -1))

    def Hsparse(self, v):                                                   # :18-27
        up = torch.cat((v[1:], v.new_zeros(1)))
        dn = torch.cat((v.new_zeros(1), v[:-1]))
        return -0.5 / self.h ** 2 * (-2 * v + up + dn) + self.potential * v

    @staticmethod
    def Hadjoint_to_padjoint(v1, v2):                                       # :29-34
        return v1 * v2

    def loss_dense(self, k: int, draw):                                     # :53-63
        Dom, _ = dense_dominant_symeig(draw)
        H = self.kinetic_dense() + torch.diag(self.potential)
        _, psi0 = Dom.apply(H, k)
        return 1.0 - (psi0.abs() * self.target).sum()

    def loss_sparse(self, k: int, draw):                                    # :64-73
        Dom, _ = make_sparse_primitives(self.Hsparse, self.Hadjoint_to_padjoint, draw)
        _, psi0 = Dom.apply(self.potential, k, self.N)
        return 1.0 - (psi0.abs() * self.target).sum()


# ----------------------------------------------------------------------------------------------
# Non-symmetric family (eig.py).  The arithmetic of the reference lives in scipy (ARPACK `eigs`,
# `gmres`; scipy is un-pinned in the reference's requirements.txt:2 — 1.18.1 in this image), so the
# restatement keeps exactly those call sites and the reference's normalisation / adjoint formulas.
# ----------------------------------------------------------------------------------------------
def dominant_eig_triple(A: np.ndarray, k: int, which: str = "LM"):
    """(eigval, left, right) with l.r = 1 and |r| = 1.  Restates eig.py:28-39."""
    from scipy.sparse import linalg as sla
    wr, vr = sla.eigs(A, k=1, which=which, ncv=k)                         # eig.py:29
    wl, vl = sla.eigs(A.T, k=1, which=which, ncv=k)                       # :30
    assert np.allclose(wr.imag, 0.0), "the desired eigenvalue must be real"   # :31
    r = vr[:, 0].real
    l = vl[:, 0].real
    l = l / np.dot(l, r)                                                  # :36
    return wr.real, l, r


def dominant_eig_backward(A: np.ndarray, eigval, l, r, grad_eigval, grad_l, grad_r):
    """grad_A for the triple above.  Restates eig.py:45-60 (two GMRES solves, three outer products)."""
    from scipy.sparse import linalg as sla
    n = A.shape[0]
    Ap = A - eigval * np.eye(n)
    b = grad_l - r * np.dot(l, grad_l)                                    # eig.py:53
    lam_l, _ = sla.gmres(Ap, b, rtol=1e-12, atol=1e-12)                   # :54 (tol= renamed rtol= in scipy>=1.12)
    Ap = A.T - eigval * np.eye(n)
    b = grad_r - l * np.dot(r, grad_r)                                    # :56
    lam_r, _ = sla.gmres(Ap, b, rtol=1e-12, atol=1e-12)                   # :57
    return grad_eigval * l[:, None] * r - l[:, None] * lam_l - lam_r[:, None] * r   # :58-60


def mps_transfer_matrix(D: int, d: int, seed: int) -> np.ndarray:
    """The D^2 x D^2 'Gong' transfer matrix of a random real MPS tensor (test_gradient.py:6-9)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, D, D))
    return np.einsum("kij,kmn->imjn", A, A).reshape(D * D, D * D)
