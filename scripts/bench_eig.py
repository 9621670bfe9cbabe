"""Timing line for the non-symmetric primitives (eig.py family) at the VUMPS size D^2 = 10^4.

    python scripts/bench_eig.py [--D 100] [--k 40] [--reps 3]

Times DominantEig forward (two restarted Arnoldi(k) solves, A and A^T, through libdsea's dense GEMV + fused
Gram-Schmidt kernels) and backward (two restarted GMRES solves + rank-2 adjoint) on the transfer matrix of a random
MPS, next to scipy's eigs / gmres (what the reference's eig.py:29-30,54,57 calls) on the host, and reports the dense
GEMV's achieved bandwidth (8 n^2 bytes per application).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dominantsparseeigenad_b200 as dsea  # noqa: E402
from dominantsparseeigenad_b200 import eig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=100)
    ap.add_argument("--k", type=int, default=40)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-scipy", action="store_true")
    a = ap.parse_args()
    rt = dsea.runtime.context()
    D, n = a.D, a.D * a.D
    gen = torch.Generator().manual_seed(1)
    A3 = torch.randn(2, D, D, dtype=torch.float64, generator=gen).cuda()
    T = torch.einsum("kij,kmn->imjn", A3, A3).reshape(n, n).contiguous()

    def sync_time(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t, out

    fwd, bwd = [], []
    for rep in range(a.reps + 1):
        Tg = T.clone().requires_grad_(True)
        l0 = rt.launch_count()
        tf, (lam, l, r) = sync_time(lambda: eig.DominantEig.apply(Tg, a.k))
        loss = lam.sum() + (l * torch.arange(n, device="cuda", dtype=torch.float64)).sum() * 1e-3 + r.sum() * 1e-3
        tb, _ = sync_time(lambda: loss.backward())
        launches = rt.launch_count() - l0
        if rep:                                   # first call warms up allocator / module loading
            fwd.append(tf)
            bwd.append(tb)
    # dense GEMV bandwidth
    op = dsea.DenseOperator(T)
    v = torch.randn(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        op.matvec_raw(None, v)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        op.matvec_raw(None, v)
    e1.record()
    torch.cuda.synchronize()
    gemv_ms = e0.elapsed_time(e1) / 20
    out = {"workload": f"DominantEig on a dense {n}x{n} MPS transfer matrix (D={D}), k={a.k}",
           "forward_s": float(np.median(fwd)), "backward_s": float(np.median(bwd)), "libdsea_launches": launches,
           "eigval": float(lam.item()), "dense_gemv_ms": gemv_ms, "dense_gemv_GBps": 8.0 * n * n / gemv_ms / 1e6,
           "residual_right": float((T @ r - lam * r).norm().item()),
           "residual_left": float((T.T @ l - lam * l).norm().item() / l.norm().item())}
    if not a.no_scipy:
        import scipy.sparse.linalg as sla
        Tn = T.cpu().numpy()
        t = time.perf_counter()
        w, vr = sla.eigs(Tn, k=1, which="LM", ncv=a.k)
        w2, vl = sla.eigs(Tn.T, k=1, which="LM", ncv=a.k)
        out["scipy_eigs_forward_s"] = time.perf_counter() - t
        out["scipy_eigval"] = float(w[0].real)
        out["host_cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
