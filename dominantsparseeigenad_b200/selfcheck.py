"""Self-check of the (optionally sharded) TFIM path through the public API, against exact identities.

`bench.py` runs this inside every multi-GPU bench so that the sharded path's parity is observed by whoever
runs the bench (the `-m gpu` multi-GPU tests skip on a 1-GPU box); `tests/mgpu_worker.py` runs it as well,
next to its element-wise comparison with the CPU oracle.  Nothing here needs the oracle:

  * character vectors chi_b[s] = (-1)^{bit_b(s)} are eigenvectors of every sigma^x_i, so
        <chi_b, (dH/dg) chi_b> = -(N - 2) 2^N          for EVERY spin bit b — local, strided and remote;
  * <1, H 1> = -g N 2^N  (the diagonal sums to zero for N > 2);
  * H is symmetric: <w, H v> = <H w, v>; the adjoint contraction equals <v1, (dH/dg) v2>;
  * a full solve at a small size reproduces the closed-form E0, dE0/dg, d2E0/dg2 and chi_F
    (`analytic.tfim_exact`), |psi0| = 1 through the shard-aware dot, replicated scalars identical on all ranks.
"""
from __future__ import annotations

import torch

from . import runtime
from .analytic import tfim_exact


def _rel(a: float, b: float) -> float:
    return abs(a - b) / max(abs(b), 1e-300)


def operator_identities(model, g: float = 1.25) -> dict:
    """Identity checks of H, dH/dg and the adjoint contraction on `model` (a TFIM of any size / sharding)."""
    from .operators import dot
    rt = model.rt
    dev, N, n_loc = model.device, model.N, model.n_loc
    L = N - rt.log2world
    s = torch.arange(n_loc, device=dev, dtype=torch.int64) | (rt.rank << L)
    model.g = torch.tensor([g], dtype=torch.float64, device=dev)
    dim = float(model.dim)
    out = {}
    worst = 0.0
    for b in range(N):
        chi = (1.0 - 2.0 * ((s >> b) & 1)).to(torch.float64)
        val = dot(chi, model.pHpg(chi)).item()
        worst = max(worst, _rel(val, -(N - 2) * dim))
    out["dHdg_character_vectors_max_rel_err"] = worst
    ones = torch.ones(n_loc, dtype=torch.float64, device=dev)
    out["ones_H_ones_rel_err"] = _rel(dot(ones, model.H(ones)).item(), -g * N * dim)
    del s, ones
    v = runtime.start_vector(n_loc, "check")
    w = runtime.start_vector(n_loc, "check")
    a, b_ = dot(w, model.H(v)).item(), dot(model.H(w), v).item()
    out["H_symmetry_rel_err"] = abs(a - b_) / max(abs(a), abs(b_), 1e-300)
    adj = model.Hadjoint_to_gadjoint(w, v).item()
    ref = dot(w, model.pHpg(v)).item()
    out["adjoint_vs_dHdg_rel_err"] = abs(adj - ref) / max(abs(ref), 1e-300)
    return out


def small_solve(N: int, k: int = 160, g: float = 1.25) -> dict:
    """E0, dE0, d2E0, chi_F of a small chain through DominantSparseSymeig vs the closed forms."""
    from . import symeig
    from .operators import TFIM, dot
    model = TFIM(N)
    dev = model.device
    prev = symeig.DominantSparseSymeig
    prim = symeig.setDominantSparseSymeig(model.H, model.Hadjoint_to_gadjoint)
    ex = tfim_exact(N, g)
    model.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
    E0, psi0 = prim.apply(model.g, k, model.dim, dev)
    dE0, = torch.autograd.grad(E0, model.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, model.g)
    norm = dot(psi0.detach(), psi0.detach()).item()
    model.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
    E0b, psib = prim.apply(model.g, k, model.dim, dev)
    logF = torch.log(dot(psib.detach(), psib))                      # chiF.py:49 with the shard-aware dot
    dlogF, = torch.autograd.grad(logF, model.g, create_graph=True)
    d2logF, = torch.autograd.grad(dlogF, model.g)
    out = {"N": N, "k": k, "g": g,
           "E0_rel_err": _rel(E0.item(), ex.E0), "dE0_rel_err": _rel(dE0.item(), ex.dE0),
           "d2E0_rel_err": _rel(d2E0.item(), ex.d2E0), "chiF_rel_err": _rel(-d2logF.item(), ex.chiF),
           "psi_norm_err": abs(norm - 1.0)}
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        t = torch.tensor([E0.item(), dE0.item(), d2E0.item(), d2logF.item()], dtype=torch.float64, device=dev)
        lo, hi = t.clone(), t.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        out["replicated_scalars_identical"] = bool(torch.equal(lo, hi))
    symeig.DominantSparseSymeig = prev
    return out


TOL = {"dHdg_character_vectors_max_rel_err": 1e-13, "ones_H_ones_rel_err": 1e-13, "H_symmetry_rel_err": 1e-11,
       "adjoint_vs_dHdg_rel_err": 1e-11, "E0_rel_err": 1e-10, "dE0_rel_err": 1e-6, "d2E0_rel_err": 1e-6,
       "chiF_rel_err": 1e-6, "psi_norm_err": 1e-12}


def verdict(result: dict) -> bool:
    ok = True
    for key, val in result.items():
        if isinstance(val, dict):
            ok = verdict(val) and ok
        elif key in TOL:
            ok = ok and (val <= TOL[key])
        elif key == "replicated_scalars_identical":
            ok = ok and bool(val)
    return ok


def run(model=None, small_N: int | None = None) -> dict:
    """Identities on `model` (or a fresh N = 16 + log2(world) chain) plus a small full solve."""
    from .operators import TFIM
    rt = runtime.context()
    if small_N is None:
        small_N = 14 + rt.log2world
    res = {"world": rt.world, "p2p": rt.p2p_enabled()}
    res["identities"] = operator_identities(model if model is not None else TFIM(16 + rt.log2world))
    res["small_solve"] = small_solve(small_N)
    res["ok"] = verdict(res)
    return res
