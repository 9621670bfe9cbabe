"""BASELINE config 1 (examples/schrodinger1D.py as shipped: N = 300 grid points, k = 300, loss = 1 - <|psi0|, target>):
one forward + backward through each variant on the GPU, checked against the seed-independent known answers
(loss 0.099454537767, |grad|_2 = 2.269836e-3), next to the unmodified reference's DominantSymeig on the CPU.

    python scripts/bench_config1.py [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dominantsparseeigenad_b200 as dsea  # noqa: E402
import dominantsparseeigenad_b200.symeig as symeig  # noqa: E402


def problem(N=300):
    x = np.linspace(-1.0, 1.0, num=N, endpoint=False)
    h = 2.0 / N
    t = np.zeros(N)
    idx = np.abs(x) < 0.5
    t[idx] = 1.0 - np.abs(x[idx])
    target = torch.from_numpy(t / np.linalg.norm(t))
    K = sp.diags([np.ones(N - 1), -2 * np.ones(N), np.ones(N - 1)], [-1, 0, 1], format="csr") * (-0.5 / h ** 2)
    return x, h, target, K


def timed(fn, reps, sync):
    fn()
    out = []
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        res = fn()
        sync()
        out.append(time.perf_counter() - t0)
    return float(np.median(out)), res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    N = k = 300
    x, h, target, K = problem(N)
    rt = dsea.runtime.context()
    out = {"workload": "schrodinger1D N=300 k=300: one forward + backward (loss, dloss/dV)"}

    def run(variant):
        dev = torch.device("cuda") if variant != "callback_cpu" else torch.device("cpu")
        V = (0.5 * torch.from_numpy(x) ** 2).to(dev).requires_grad_(True)
        tgt = target.to(dev)
        if variant == "dense":
            Kd = torch.from_numpy(K.toarray()).to(dev)
            _, psi = symeig.DominantSymeig.apply(Kd + torch.diag(V), k)
        elif variant == "csr":
            op = dsea.SparseMatrixOperator.from_scipy(K, V)
            symeig.setDominantSparseSymeig(op.H, op.Hadjoint_to_padjoint)
            _, psi = symeig.DominantSparseSymeig.apply(V, k, N, dev)
        else:                                                   # the reference's closures on CPU tensors, as shipped
            zero = torch.zeros(1, dtype=torch.float64)
            Hs = lambda v: -0.5 / h ** 2 * (-2 * v + torch.cat((v[1:], zero)) + torch.cat((zero, v[:-1]))) + V * v
            symeig.setDominantSparseSymeig(Hs, lambda v1, v2: v1 * v2)
            _, psi = symeig.DominantSparseSymeig.apply(V, k, N)
        loss = 1.0 - (psi.abs() * tgt).sum()
        g, = torch.autograd.grad(loss, V)
        return loss.item(), g.norm().item()

    for variant in ("dense", "csr", "callback_cpu"):
        t, (loss, gn) = timed(lambda: run(variant), a.reps, torch.cuda.synchronize)
        out[variant] = {"seconds_fwd_bwd": t, "loss": loss, "grad_norm": gn,
                        "loss_err": abs(loss - 0.099454537767), "grad_norm_rel_err": abs(gn - 2.269836e-3) / 2.269836e-3}
    try:
        from baseline import ref_runner
        ref_symeig, _ = ref_runner.load()
        torch.set_num_threads(os.cpu_count() or 1)
        Kd = torch.from_numpy(K.toarray())

        def ref():
            V = (0.5 * torch.from_numpy(x) ** 2).requires_grad_(True)
            _, psi = ref_symeig.DominantSymeig.apply(Kd + torch.diag(V), k)
            loss = 1.0 - (psi.abs() * target).sum()
            g, = torch.autograd.grad(loss, V)
            return loss.item(), g.norm().item()

        t, (loss, gn) = timed(ref, a.reps, lambda: None)
        out["reference_cpu_dense"] = {"seconds_fwd_bwd": t, "loss": loss, "grad_norm": gn, "cores": os.cpu_count()}
    except Exception as exc:
        out["reference_cpu_dense"] = {"unavailable": str(exc)[:200]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
