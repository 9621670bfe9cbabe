"""Micro-benchmark of the TFIM operator kernels (K1 / K6) on one GPU.

    python scripts/bench_matvec.py [--spins 24 26] [--reps 20] [--out file.json]

For every size it times dsea_matvec (plain, and with the shift + dot epilogue the CG loop uses) and
dsea_adjoint with CUDA events under each kernel variant, checks that all variants agree with the first one,
and prints achieved GB/s of the ALGORITHMIC 16 n bytes next to the HBM bytes the sweep plan actually moves
(16 n first sweep + 24 n per further sweep).

Variants (dsea_ctx_set_option knobs):
  r1_plan      round 1's plan: 32-byte runs (run_bits = 2), no direct bits, 512 threads
  sweeps3      128-byte runs, no direct bits (a third sweep for the top bits)
  direct       128-byte runs, top local bits by direct (L2-served) loads                  <- default build
  staged       ... strided sweep with its epilogue operands staged through thread-private shared-memory slots   <- default build
  direct_allpipe / direct_lastgeneric  ... last sweep always pipelined / always the generic 2-CTA/SM kernel
  direct_pf    ... with prefetch.global.L2 hints for the next tile's epilogue operands
  direct_nounroll ... flip-bit loop bounds taken at run time (no unrolling)
  direct_256 / direct_1024  ... 256 threads x 16 pairs / 1024 threads x 4 pairs (5 / 3 register-resident tile bits)
  sweeps3_1024 three sweeps with 1024 threads
  direct_notma ... contiguous tiles staged with LDGSTS instead of TMA bulk copies
  generic      the non-pipelined 2-CTA/SM kernel with the default plan
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dominantsparseeigenad_b200 as dsea  # noqa: E402
from dominantsparseeigenad_b200 import _lib  # noqa: E402
from dominantsparseeigenad_b200.runtime import ptr, stream_ptr  # noqa: E402

DEFAULTS = {"tfim_pipeline": 1, "tfim_tma": 1, "tfim_run_bits": 0, "tfim_direct": 1, "tfim_pipe_threads": 512,
            "tfim_l2_prefetch": 0, "tfim_pipe_adjoint": 1, "tfim_unroll": 1, "tfim_generic_min_operands": 4,
            "tfim_stage": 0}
VARIANTS = {
    "r1_plan": {"tfim_run_bits": 2, "tfim_direct": 0, "tfim_unroll": 0},
    "sweeps3": {"tfim_direct": 0},
    "direct": {},
    "staged": {"tfim_stage": 1},
    "direct_allpipe": {"tfim_generic_min_operands": 99},
    "direct_lastgeneric": {"tfim_generic_min_operands": 1},
    "direct_pf": {"tfim_l2_prefetch": 1},
    "direct_nounroll": {"tfim_unroll": 0},
    "direct_256": {"tfim_pipe_threads": 256},
    "direct_1024": {"tfim_pipe_threads": 1024},
    "sweeps3_1024": {"tfim_direct": 0, "tfim_pipe_threads": 1024},
    "direct_notma": {"tfim_tma": 0},
    "generic": {"tfim_pipeline": 0},
}


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
    ap.add_argument("--variants", nargs="+", default=list(VARIANTS))
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rt = dsea.runtime.context()
    lib = rt.lib
    peak = 6540.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
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
        adj = torch.empty(1, dtype=torch.float64, device="cuda")
        ref = None
        for name in args.variants:
            for key, val in {**DEFAULTS, **VARIANTS[name]}.items():
                rt.set_option(key, val)
            buf = (C.c_int * 200)()
            ns = lib.dsea_tfim_plan(N, 13, {**DEFAULTS, **VARIANTS[name]}["tfim_run_bits"],
                                    {**DEFAULTS, **VARIANTS[name]}["tfim_direct"], buf)
            plan = [tuple(buf[5 * j:5 * j + 5]) for j in range(ns)]
            st = stream_ptr()
            plain = lambda: _lib.check(lib.dsea_matvec(rt.handle, m.handle, ptr(g), None, ptr(v), ptr(u), None, None, st))
            cgmv = lambda: _lib.check(lib.dsea_matvec(rt.handle, m.handle, ptr(g), ptr(sh), ptr(v), ptr(u), ptr(dot), None, st))
            adjf = lambda: _lib.check(lib.dsea_adjoint(rt.handle, m.handle, ptr(w), ptr(v), ptr(adj), None, st))
            plain()
            torch.cuda.synchronize()
            got = u.clone()
            adjf()
            a_val = adj.item()
            if ref is None:
                ref = (got, a_val)
            err = (got - ref[0]).abs().max().item() / ref[0].abs().max().item()
            aerr = abs(a_val - ref[1]) / abs(ref[1])
            assert err < 1e-13 and aerr < 1e-11, (name, err, aerr)
            t_plain, t_cg, t_adj = time_ms(plain, args.reps), time_ms(cgmv, args.reps), time_ms(adjf, args.reps)
            hbm_bytes = (16 + 24 * (ns - 1)) * n
            row = {"N": N, "variant": name, "plan": plan, "matvec_ms": t_plain, "matvec_shift_dot_ms": t_cg,
                   "adjoint_ms": t_adj, "algorithmic_GBs": 16 * n / t_plain / 1e6,
                   "frac_of_16n_roofline": 16 * n / t_plain / 1e6 / peak,
                   "plan_hbm_bytes_per_el": hbm_bytes / n, "frac_of_plan_traffic": hbm_bytes / t_plain / 1e6 / peak,
                   "max_rel_diff_vs_first_variant": err}
            out.append(row)
            print(json.dumps(row), flush=True)
        for key, val in DEFAULTS.items():
            rt.set_option(key, val)
        del m, v, w, u
        torch.cuda.empty_cache()
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
