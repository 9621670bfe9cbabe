"""Micro-benchmark of the TFIM operator kernels (K1 / K6) on one GPU.

    python scripts/bench_matvec.py [--spins 24 26] [--reps 20]

For every size it times dsea_matvec (with and without the dot epilogue + shift) and dsea_adjoint with
CUDA events, for the generic sweep kernel and for the persistent double-buffered one, checks that the two
agree, and prints achieved GB/s of the ALGORITHMIC 16 n bytes next to the structural bounds of the
two-sweep scheme (40 n bytes of HBM traffic; 8 N n bytes through the shared-memory crossbar).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dominantsparseeigenad_b200 as dsea  # noqa: E402
from dominantsparseeigenad_b200 import _lib  # noqa: E402
from dominantsparseeigenad_b200.runtime import ptr, stream_ptr  # noqa: E402


def time_ms(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spins", type=int, nargs="+", default=[24])
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rt = dsea.runtime.context()
    lib = rt.lib
    out = []
    for N in args.spins:
        m = dsea.TFIM(N)
        n = m.n_loc
        g = torch.tensor([1.1], dtype=torch.float64, device="cuda")
        sh = torch.tensor([-3.0], dtype=torch.float64, device="cuda")
        v = torch.randn(n, dtype=torch.float64, device="cuda")
        w = torch.randn(n, dtype=torch.float64, device="cuda")
        u = torch.empty_like(v)
        dot = torch.empty(1, dtype=torch.float64, device="cuda")
        res = {}
        ref = None
        for pipe, tma in ((0, 0), (1, 0), (1, 1)):
            rt.set_option("tfim_pipeline", pipe)
            rt.set_option("tfim_tma", tma)
            st = stream_ptr()
            plain = lambda: _lib.check(lib.dsea_matvec(rt.handle, m.handle, ptr(g), None, ptr(v), ptr(u), None, None, st))
            cgmv = lambda: _lib.check(lib.dsea_matvec(rt.handle, m.handle, ptr(g), ptr(sh), ptr(v), ptr(u), dot.data_ptr(), None, st))
            adj = lambda: _lib.check(lib.dsea_adjoint(rt.handle, m.handle, ptr(w), ptr(v), dot.data_ptr(), None, st))
            t_plain, t_cg, t_adj = time_ms(plain, args.reps), time_ms(cgmv, args.reps), time_ms(adj, args.reps)
            plain()
            torch.cuda.synchronize()
            if ref is None:
                ref = u.clone()
                err = 0.0
            else:
                err = (u - ref).abs().max().item() / ref.abs().max().item()
            adj()
            adjval = dot.item()
            res[f"pipeline{pipe}_tma{tma}"] = {"matvec_ms": t_plain, "matvec_GBps_of_16n": 16 * n / t_plain / 1e6,
                                      "matvec_cg_ms": t_cg, "adjoint_ms": t_adj,
                                      "adjoint_GBps_of_16n": 16 * n / t_adj / 1e6, "rel_diff_vs_generic": err,
                                      "adjoint_value": adjval}
        rt.set_option("tfim_pipeline", 1)
        rt.set_option("tfim_tma", 0)
        res["bounds_ms_at_6540GBps"] = {"algorithmic_16n": 16 * n / 6540e6, "two_sweep_hbm_40n": 40 * n / 6540e6}
        out.append({"spins": N, **res})
        del m, v, w, u, ref
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
