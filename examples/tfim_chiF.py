"""Fidelity susceptibility chi_F(g) of the 1-D TFIM by second-order AD through the dominant eigensolver.

Same experiment as the reference's examples/TFIM/chiF.py (chiF.py:40-82).  Under torchrun the state is
sharded over the GPUs; the only change to the user code is `dsea.dot` instead of `matmul` for the global
inner product (chiF.py:49).

    python examples/tfim_chiF.py --spins 20 --k 100 --points 5
    torchrun --nproc-per-node 8 examples/tfim_chiF.py --spins 28 --k 200 --points 1
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sweep(N, k, gs, save=""):
    """chiF.py:65-82: chi_F over a list of g values; returns rows (g, chi_F) (rank-replicated under torchrun)."""
    import dominantsparseeigenad_b200 as dsea
    import dominantsparseeigenad_b200.symeig as symeig
    rt = dsea.runtime.context()
    model = dsea.TFIM(N)
    rows = []
    for gv in gs:
        model.g = torch.tensor([float(gv)], dtype=torch.float64, device=model.device, requires_grad=True)
        symeig.setDominantSparseSymeig(model.H, model.Hadjoint_to_gadjoint)                         # chiF.py:46
        E0, psi0 = symeig.DominantSparseSymeig.apply(model.g, min(k, model.dim), model.dim, model.device)
        logF = torch.log(dsea.dot(psi0.detach(), psi0))                                             # chiF.py:49
        dlogF, = torch.autograd.grad(logF, model.g, create_graph=True)
        d2logF, = torch.autograd.grad(dlogF, model.g)
        rows.append((float(gv), -d2logF.item()))
    if save and rt.rank == 0:
        arr = np.array(rows)
        np.savez(save, gs=arr[:, 0], chiFs=arr[:, 1])                                               # chiF.py:81-82
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spins", type=int, default=16)
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--points", type=int, default=5)
    ap.add_argument("--gmin", type=float, default=1.0)
    ap.add_argument("--gmax", type=float, default=1.5)
    ap.add_argument("--save", default="")
    args = ap.parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    import dominantsparseeigenad_b200 as dsea
    import dominantsparseeigenad_b200.symeig as symeig
    rt = dsea.runtime.context()
    model = dsea.TFIM(args.spins)
    gs = np.linspace(args.gmin, args.gmax, num=args.points)
    chis = np.empty(args.points)
    for i, gv in enumerate(gs):
        model.g = torch.tensor([gv], dtype=torch.float64, device=model.device, requires_grad=True)
        symeig.setDominantSparseSymeig(model.H, model.Hadjoint_to_gadjoint)                         # chiF.py:46
        E0, psi0 = symeig.DominantSparseSymeig.apply(model.g, args.k, model.dim, model.device)
        logF = torch.log(dsea.dot(psi0.detach(), psi0))                                             # chiF.py:49
        dlogF, = torch.autograd.grad(logF, model.g, create_graph=True)
        d2logF, = torch.autograd.grad(dlogF, model.g)
        chis[i] = -d2logF.item()
        if rt.rank == 0:
            print(f"g = {gv:.6f}   E0/N = {E0.item() / args.spins:+.12f}   chi_F = {chis[i]:.10f}")
    if args.save and rt.rank == 0:
        np.savez(args.save, gs=gs, chiFs=chis)                                                      # chiF.py:81-82


if __name__ == "__main__":
    main()
