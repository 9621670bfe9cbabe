"""Worker for the multi-GPU parity test: launched by torchrun (one process per GPU, NCCL).

Checks the sharded path (state vector split by its top log2(world) spin bits, pairwise NCCL exchange of
shards for the top-bit flips, NCCL allreduce of every dot product) against the CPU oracle on the same
seeded global vectors, and the E0 / dE0 / d2E0 / chi_F family against the analytic values.
Prints 'MGPU_OK' on rank 0 when everything passed.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    import dominantsparseeigenad_b200 as dsea
    from oracle import dsea_oracle as orc
    rt = dsea.runtime.context()
    assert rt.world == world and rt.rank == rank
    dev = rt.device
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)

    # ---- matvec / dH/dg / adjoint / dot against the oracle on the same global vectors
    for N, tile_bits in ((12, 13), (14, 6), (16, 13)):
        rt.set_option("tfim_tile_bits", tile_bits)
        g = 0.9 + 0.01 * N
        n, n_loc = 1 << N, (1 << N) // world
        rng = np.random.default_rng(N)
        v, w = rng.standard_normal(n), rng.standard_normal(n)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        m = dsea.TFIM(N)
        assert m.n_loc == n_loc and m.dim == n
        m.g = torch.tensor([g], dtype=torch.float64, device=dev)
        o = orc.TFIMOracle(N, g)
        vl, wl = torch.from_numpy(v[sl]).to(dev), torch.from_numpy(w[sl]).to(dev)
        want = o.H(torch.from_numpy(v)).numpy()
        got = m.H(vl).cpu().numpy()
        assert np.abs(got - want[sl]).max() <= 1e-14 * N * np.abs(want).max(), ("H", N, rank)
        wantp = o.pHpg(torch.from_numpy(v)).numpy()
        assert np.abs(m.pHpg(vl).cpu().numpy() - wantp[sl]).max() <= 1e-13 * N * np.abs(v).max(), ("pHpg", N, rank)
        want_adj = o.Hadjoint_to_gadjoint(torch.from_numpy(w), torch.from_numpy(v)).item()
        assert rel(m.Hadjoint_to_gadjoint(wl, vl).item(), want_adj) < 1e-11, ("adj", N, rank)
        assert rel(dsea.dot(vl, wl).item(), float(np.dot(v, w))) < 1e-11, ("dot", N, rank)
    rt.set_option("tfim_tile_bits", 13)

    # ---- the sharded random start vector equals the single-GPU one (counter = global index)
    torch.manual_seed(77)
    dsea.runtime.reset_draw_counter()
    mine = dsea.runtime.start_vector(1 << 10, "lanczos").cpu()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather_object(gathered, mine)
    full = torch.cat(gathered)
    assert abs(full.mean().item()) < 0.15 and abs(full.std().item() - 1) < 0.1

    # ---- E0, dE0, d2E0, chiF at N=16 (k=200) against analytic values; eigenvector norm via dist dot
    N, k, g = 16, 200, 1.25
    m = dsea.TFIM(N)
    m.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
    E0, psi0 = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, dev)
    dE0, = torch.autograd.grad(E0, m.g, create_graph=True)
    d2E0, = torch.autograd.grad(dE0, m.g)
    aE0, adE0, ad2E0, achi = orc.tfim_analytic(N, g)
    assert psi0.shape == (m.n_loc,)
    assert rel(E0.item(), aE0) < 1e-10, (E0.item(), aE0)
    assert rel(dE0.item(), adE0) < 1e-6, (dE0.item(), adE0)
    assert rel(d2E0.item(), ad2E0) < 1e-6, (d2E0.item(), ad2E0)
    assert abs(dsea.dot(psi0.detach(), psi0.detach()).item() - 1.0) < 1e-12
    m.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
    E0, psi0 = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, dev)
    logF = torch.log(dsea.dot(psi0.detach(), psi0))
    dlogF, = torch.autograd.grad(logF, m.g, create_graph=True)
    d2logF, = torch.autograd.grad(dlogF, m.g)
    if os.environ.get("DSEA_DIAG"):
        for rep in range(3):
            m.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
            E0b, psib = dsea.symeig.DominantSparseSymeig.apply(m.g, k, m.dim, dev)
            lf = torch.log(dsea.dot(psib.detach(), psib))
            d1, = torch.autograd.grad(lf, m.g, create_graph=True)
            d2, = torch.autograd.grad(d1, m.g)
            if rank == 0:
                print("DIAG chiF", world, rt.p2p_enabled(), -d2.item(), achi, d1.item(),
                      dsea.runtime.stats["cg_iters"][-4:], flush=True)
    assert rel(-d2logF.item(), achi) < 1e-6, (-d2logF.item(), achi)
    # opt-in fp32 shadow basis on the sharded path (remote shards are rounded exactly like the local vector)
    dsea.runtime.set_basis_precision("fp32")
    m32 = dsea.TFIM(N)
    m32.g = torch.tensor([g], dtype=torch.float64, device=dev, requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(m32.H, m32.Hadjoint_to_gadjoint)
    E32, psi32 = dsea.symeig.DominantSparseSymeig.apply(m32.g, k, m32.dim, dev)
    dE32, = torch.autograd.grad(E32, m32.g)
    dsea.runtime.set_basis_precision("fp64")
    assert rel(E32.item(), aE0) < 1e-10 and rel(dE32.item(), adE0) < 1e-6, (E32.item(), aE0, dE32.item(), adE0)
    assert 1 - abs(dsea.dot(psi32.detach(), psi0.detach()).item()) < 1e-8
    # opt-in even-parity start vectors: the sharded spin flip pairs rank r with rank P-1-r
    dsea.runtime.parity_sector = "even"
    torch.manual_seed(5)
    v_even = dsea.runtime.start_vector(m.n_loc, "lanczos")
    assert (v_even - dsea.runtime.spin_flip(v_even)).abs().max().item() < 1e-12
    mp = dsea.TFIM(N)
    mp.g = torch.tensor([0.5], dtype=torch.float64, device=dev, requires_grad=True)
    dsea.symeig.setDominantSparseSymeig(mp.H, mp.Hadjoint_to_gadjoint)
    Ep, psip = dsea.symeig.DominantSparseSymeig.apply(mp.g, k, mp.dim, dev)
    dEp, = torch.autograd.grad(Ep, mp.g, create_graph=True)
    d2Ep, = torch.autograd.grad(dEp, mp.g)
    dsea.runtime.parity_sector = "none"
    pE0, pdE0, pd2E0, _ = orc.tfim_analytic(N, 0.5)
    assert rel(Ep.item(), pE0) < 1e-10 and rel(dEp.item(), pdE0) < 1e-6 and rel(d2Ep.item(), pd2E0) < 1e-6, (d2Ep.item(), pd2E0)
    assert (psip.detach() - dsea.runtime.spin_flip(psip.detach())).abs().max().item() < 1e-9
    dsea.symeig.setDominantSparseSymeig(m.H, m.Hadjoint_to_gadjoint)
    # the oracle-free identities bench.py asserts in every multi-GPU run (each spin bit, remote ones included)
    from dominantsparseeigenad_b200 import selfcheck
    sc = selfcheck.run()
    assert sc["ok"], sc
    # every rank holds identical replicated scalars
    t = torch.tensor([E0.item(), dE0.item(), d2E0.item()], dtype=torch.float64, device=dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} E0={E0.item():.12f} dE0={dE0.item():.9f} d2E0={d2E0.item():.9f} chiF={-d2logF.item():.9f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
