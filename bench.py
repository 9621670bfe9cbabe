#!/usr/bin/env python
"""bench.py — forward+backward seconds per dominant-eigenpair solve on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload E0_dE0|chiF]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one solve of the TFIM ground state through the reference-shaped public API:
    E0, psi0 = DominantSparseSymeig.apply(g, k, dim)         Lanczos, k vectors, full reorthogonalisation
    dE0,     = torch.autograd.grad(E0, g)                    CG solve of (H - E0) x = b + adjoint contraction
(examples/TFIM/E0.py:60-63 shape; `--workload chiF` runs examples/TFIM/chiF.py:46-53 instead: one forward and
two nested backward passes = three CG solves).  At 1 GPU the workload is TFIM N=24, k=200 — the largest
BASELINE configuration whose fp64 Lanczos basis (26.8 GB) fits one GPU; at P GPUs the state vector is sharded
by its top log2(P) spin bits with n_loc = 2^24 amplitudes per GPU (weak scaling: N = 24 + log2 P).

Prints ONE JSON line (rank 0).
  value         device-timed seconds per solve with g resident in HBM;
  e2e           the same through host buffers (g from pinned host memory, E0 / dE0 / psi0 copied back);
  roofline      the dominant kernel (re-orthogonalisation pass 2) timed live with CUDA events in the timed region;
  cpu_baseline  the UNMODIFIED reference (baseline/_ref) timed on the host cores on a bounded sample — TFIM N=20,
                k=100 (BASELINE config 2) — `value` is the MEASURED time of that sample, never an extrapolation;
  pairs         the same sample solved by this library in the same run, so a measured same-config ratio exists;
  headline      (P >= 4) TFIM N=28, k=200 (BASELINE config 4) and (P = 8) N=30, k=200 (config 5, fp32 shadow basis)
                measured in the same process after the weak-scaling line;
  selfcheck     (P > 1) parity of the sharded path against exact identities and closed forms.
`--impl reference` times only the reference's CPU path: the product configuration itself when the host can run
it inside the time budget (N <= 24), otherwise the bounded sample, labelled as what it is.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "tfim_fwd_bwd_seconds_per_solve"
UNIT = "s/solve"
SAMPLE_N, SAMPLE_K = 20, 100          # CPU sample: BASELINE config 2 shape (largest upstream fixture size)
MODES = {"E0_dE0": "E0_plus_dE0dg", "chiF": "chiF"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="E0_dE0", choices=sorted(MODES))
    ap.add_argument("--spins", type=int, default=0, help="override N (default 24 + log2 gpus)")
    ap.add_argument("--k", type=int, default=200, help="Lanczos vectors; 0 = largest k <= 200 whose basis fits HBM")
    ap.add_argument("--g", type=float, default=1.0)
    ap.add_argument("--basis", default="fp64", choices=["fp64", "fp32"],
                    help="storage of the Lanczos basis (fp32 = opt-in shadow basis with fp64 accumulation)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-2 pair / headline configs / selfcheck")
    ap.add_argument("--cpu-sample-spins", type=int, default=SAMPLE_N, help="(tests) shrink the bounded CPU sample")
    ap.add_argument("--cpu-sample-k", type=int, default=SAMPLE_K)
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("DSEA_REF_BUDGET_S", 420)),
                    help="reference arm: wall-clock budget for attempting the product configuration itself")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.fh.close()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (baseline/_ref) or, if it is not staged, the oracle port
# ------------------------------------------------------------------------------------------------
def workload_name(N, k, mode):
    return f"tfim_N{N}_k{k}_{MODES[mode]}"


def host_info():
    info = {"cores": os.cpu_count() or 1}
    try:
        import psutil
        vm = psutil.virtual_memory()
        info["ram_total_gb"], info["ram_available_gb"] = vm.total / 1e9, vm.available / 1e9
    except Exception:
        pass
    return info


def _port_solve(N, k, g, mode, seed):
    """Fallback when baseline/_ref is absent: the oracle's restatement of the same path (kind = "port")."""
    from oracle import dsea_oracle as orc
    model = orc.TFIMOracle(N, torch.tensor([g], dtype=torch.float64, requires_grad=True))
    stats = {}
    Dom, _ = orc.make_sparse_primitives(model.H, model.Hadjoint_to_gadjoint, orc.SeededDraws(seed), stats)
    t0 = time.perf_counter()
    E0, psi0 = Dom.apply(model.g, k, model.dim)
    t1 = time.perf_counter()
    if mode == "chiF":
        logF = torch.log(psi0.detach().matmul(psi0))
        d1, = torch.autograd.grad(logF, model.g, create_graph=True)
        d2, = torch.autograd.grad(d1, model.g)
        extra = {"chiF": -d2.item()}
    else:
        dE0, = torch.autograd.grad(E0, model.g)
        extra = {"dE0": dE0.item()}
    t2 = time.perf_counter()
    return {"N": N, "k": k, "g": g, "mode": mode, "fwd": t1 - t0, "bwd": t2 - t1, "total": t2 - t0, "E0": E0.item(),
            **extra}


def cpu_solve(N, k, g, mode, seed=1234):
    """(kind, timing dict) of ONE measured CPU solve at exactly (N, k, g, mode)."""
    from baseline import ref_runner
    if ref_runner.available():
        ref_runner.warm(N)
        return "reference", ref_runner.solve(N, k, g, mode, seed)
    return "port", _port_solve(N, k, g, mode, seed)


def predict_seconds(t, Ns, ks, Nw, kw):
    """Planning estimate ONLY (never reported as a measurement): scales a measured sample solve to another size
    with the reference's own cost model: operator ~ (16N+24) 2^N bytes per call, reorth and the
    Ritz GEMM ~ 2^N k^2, CG iteration count held fixed.  Used to decide whether a full-size run fits the budget."""
    rn = 2.0 ** (Nw - Ns)
    op = rn * (16 * Nw + 24) / (16 * Ns + 24)
    fwd = t.get("mv_fwd", 0.5 * t["fwd"])
    return fwd * op * (kw / ks) + (t["fwd"] - fwd) * rn * (kw / ks) ** 2 + t["bwd"] * op


def cpu_baseline_leg(args, mode):
    """Bounded sample for the product arm's `cpu_baseline`: one measured solve at N=20, k=100."""
    torch.set_num_threads(os.cpu_count() or 1)
    kind, t = cpu_solve(SAMPLE_N, SAMPLE_K, args.g, mode)
    return {"value": t["total"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample_workload": workload_name(SAMPLE_N, SAMPLE_K, mode),
            "sample": (f"ONE measured solve of TFIM N={SAMPLE_N}, k={SAMPLE_K}, g={args.g} ({MODES[mode]}) by the "
                       f"{'unmodified reference staged under baseline/_ref' if kind == 'reference' else 'oracle port'}"
                       f" on {torch.get_num_threads()} host threads after one warm-up operator application: "
                       f"fwd {t['fwd']:.2f} s + bwd {t['bwd']:.2f} s; value is that measurement, NOT scaled to the "
                       f"bench workload"),
            "detail": t}


def run_reference_arm(args, N, k, mode, config):
    """`--impl reference`: measures the reference's own CPU path.  Returns the JSON line."""
    t_start = time.perf_counter()
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    host = host_info()
    from baseline import ref_runner
    kind, sample = cpu_solve(SAMPLE_N, SAMPLE_K, args.g, mode)
    attempt = {"attempted": False}
    full = None
    need_gb = ref_runner.required_host_bytes(N, k) / 1e9
    if (N, k) == (SAMPLE_N, SAMPLE_K):
        full = sample
    elif N > 25:
        attempt["why_not"] = (f"the reference cannot run N={N}: its (2^N, N) int64 flip table alone is "
                              f"{8 * N * 2 ** N / 1e9:.0f} GB (TFIM.py:48-51)")
    elif host.get("ram_available_gb", 0.0) < 1.15 * need_gb:
        attempt["why_not"] = f"needs ~{need_gb:.0f} GB of host RAM, {host.get('ram_available_gb', 0):.0f} GB available"
    else:
        pred = predict_seconds(sample, SAMPLE_N, SAMPLE_K, N, k)
        attempt["predicted_seconds"] = pred
        left = args.ref_budget_s - (time.perf_counter() - t_start)
        if pred > 0.8 * left:
            attempt["why_not"] = f"predicted {pred:.0f} s exceeds the {args.ref_budget_s:.0f} s budget (--ref-budget-s)"
        elif not ref_runner.available():
            attempt["why_not"] = "baseline/_ref is not staged"
        else:
            attempt["attempted"] = True
            cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_runner.py"), "--spins", str(N), "--k", str(k),
                   "--g", str(args.g), "--mode", mode]
            try:       # own process: ~60 GB of host memory is returned to the OS afterwards, and it can be timed out
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=left)
                if out.returncode == 0:
                    full = json.loads(out.stdout.strip().splitlines()[-1])
                else:
                    attempt["why_not"] = "full-size run failed: " + out.stderr.strip()[-300:]
            except subprocess.TimeoutExpired:
                attempt["why_not"] = f"full-size run exceeded the remaining budget ({left:.0f} s) and was stopped"
    if full is not None:
        measured, Nm, km = full, N, k
        config = dict(config)
    else:
        measured, Nm, km = sample, SAMPLE_N, SAMPLE_K
        config = dict(config, workload=workload_name(Nm, km, mode), spins=Nm, lanczos_vectors=km,
                      sharding="host_cpu", product_workload=workload_name(N, k, mode),
                      note="bounded sample: the product workload itself was not run on the CPU (see cpu_baseline.attempt)")
        config.pop("l2", None)
    value = measured["total"]
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": (f"ONE measured solve of {workload_name(Nm, km, mode)} at g={args.g} by the "
                      f"{'unmodified reference (baseline/_ref)' if kind == 'reference' else 'oracle port'} on {cores} "
                      f"host threads: fwd {measured['fwd']:.2f} s, bwd {measured['bwd']:.2f} s"),
           "sample_workload": workload_name(Nm, km, mode), "detail": measured, "attempt": attempt, "host": host}
    if full is not None and (Nm, km) != (SAMPLE_N, SAMPLE_K):
        cpu["config2_sample"] = {"workload": workload_name(SAMPLE_N, SAMPLE_K, mode), "value": sample["total"],
                                 "detail": sample}
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_seconds": time.perf_counter() - t_start}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    global SAMPLE_N, SAMPLE_K
    args = parse()
    SAMPLE_N, SAMPLE_K = args.cpu_sample_spins, args.cpu_sample_k
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    log2w = max(world, 1).bit_length() - 1
    mode = args.workload
    N = args.spins or (24 + log2w)
    k = args.k
    bytes_per = 4 if args.basis == "fp32" else 8
    if args.impl == "ours" and k <= 0:
        # capacity policy: the basis is k * bytes_per * n_loc bytes per GPU; keep 16 fp64 vectors + 2 GB
        # for the matvec / CG work space, the peer arena and the eigenvector.
        torch.cuda.set_device(local_rank)
        free_b, _ = torch.cuda.mem_get_info()
        vec_b = 8 * 2 ** (N - log2w)
        k = int(max(2, min(200, (free_b - 16 * vec_b - 2e9) // (vec_b * bytes_per // 8))))
    elif k <= 0:
        k = 200

    def make_config(N_, k_, basis):
        return {"workload": workload_name(N_, k_, mode), "spins": N_, "lanczos_vectors": k_, "g": args.g,
                "sharding": f"top{log2w}bits_x{world}" if world > 1 else "single_gpu", "basis": basis,
                "l2": "inputs_exceed_l2 (Lanczos basis %.1f GB per GPU)"
                      % (k_ * 2.0 ** (N_ - log2w) * (4 if basis == "fp32" else 8) / 1e9),
                "cg": "as reference: eps=1e-7 absolute, random projected x0 (CG.py:25,121-122)"}

    config = make_config(N, k, args.basis)

    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(run_reference_arm(args, N, k, mode, config)))
        return

    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import dominantsparseeigenad_b200 as dsea
    from dominantsparseeigenad_b200.analytic import tfim_exact
    rt = dsea.runtime.context()
    dev = rt.device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def measure(N_, k_, basis, steps, warmup, with_e2e=True, with_clocks=True, profile=True):
        """Times `steps` solves of (N_, k_) after `warmup`; returns the fields of a bench line.  `profile`: per-kernel
        CUDA events inside the timed region (needed for the live roofline numbers; the ~1000 event records per solve
        cost 0.5 % at N=24 but 15 % at N=20, where they also break the programmatic launch chains)."""
        rt.set_option("basis_fp32", 1 if basis == "fp32" else 0)
        model = dsea.TFIM(N_)
        dsea.symeig.setDominantSparseSymeig(model.H, model.Hadjoint_to_gadjoint)
        solve = dsea.symeig.DominantSparseSymeig.apply
        n_loc = model.n_loc
        g_host = torch.tensor([args.g], dtype=torch.float64).pin_memory()
        out_host = torch.empty(2, dtype=torch.float64).pin_memory()
        psi_host = torch.empty(n_loc, dtype=torch.float64).pin_memory() if with_e2e else None
        g_dev = g_host.to(dev)
        res = {}

        def backward(E0, psi0):
            if mode == "chiF":                                           # chiF.py:49-52 (shard-aware dot)
                logF = torch.log(dsea.dot(psi0.detach(), psi0))
                d1, = torch.autograd.grad(logF, model.g, create_graph=True)
                d2, = torch.autograd.grad(d1, model.g)
                return -d2
            dE0, = torch.autograd.grad(E0, model.g)                      # E0.py:63
            return dE0

        def step_device():
            model.g = g_dev.detach().requires_grad_(True)
            E0, psi0 = solve(model.g, k_, model.dim, dev)
            res["E0"], res["grad"], res["psi0"] = E0, backward(E0, psi0), psi0

        def step_e2e():
            model.g = g_host.to(dev, non_blocking=True).requires_grad_(True)
            E0, psi0 = solve(model.g, k_, model.dim, dev)
            grad = backward(E0, psi0)
            out_host[0:1].copy_(E0.detach().reshape(1), non_blocking=True)
            out_host[1:2].copy_(grad.detach().reshape(1), non_blocking=True)
            psi_host.copy_(psi0.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(warmup):
            step_device()
        barrier()
        if profile:
            rt.profile_enable(True)
        rt.profile_collect()
        dsea.runtime.stats["cg_iters"].clear()
        l0 = rt.launch_count()
        if with_clocks:
            with ClockSampler(local_rank) as clk:
                ms_total = timed(step_device, steps)
            clocks = clk.summary()
        else:
            ms_total, clocks = timed(step_device, steps), None
        launches = rt.launch_count() - l0
        prof = rt.profile_collect()
        rt.profile_enable(False)
        out = {"N": N_, "k": k_, "n_loc": n_loc, "ms_total": ms_total, "steps": steps, "launches": launches,
               "prof": prof, "clocks": clocks, "cg_iters": list(dsea.runtime.stats["cg_iters"]),
               "E0": res["E0"].item(), "grad": res["grad"].item(),
               "psi_norm_err": abs(dsea.dot(res["psi0"].detach(), res["psi0"].detach()).item() - 1.0)}
        if with_e2e:
            step_e2e()
            out["e2e_s"] = timed(step_e2e, steps) / 1e3 / steps
        res.clear()
        del model
        rt.set_option("basis_fp32", 0)
        torch.cuda.empty_cache()
        return out

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def kernel_table(m):
        prof, steps, ms_total = m["prof"], m["steps"], m["ms_total"]
        table = {}
        for name, p in prof.items():
            if p["launches"] == 0 or p["ms"] <= 0:
                continue
            row = {"launches_per_step": p["launches"] / steps, "ms_per_step": p["ms"] / steps,
                   "share_of_step": p["ms"] / ms_total}
            if p["bytes"] > 0:
                gbs = p["bytes"] / (p["ms"] * 1e-3) / 1e9
                row.update(achieved=gbs, frac=gbs / peak, algorithmic_bytes_per_step=p["bytes"] / steps)
            table[name] = row
        table["unattributed_share_of_step"] = 1.0 - sum(r["share_of_step"] for r in table.values())
        return table

    def check(m):
        ex = tfim_exact(m["N"], args.g)
        want = ex.chiF if mode == "chiF" else ex.dE0
        return {"E0_rel_err": abs(m["E0"] - ex.E0) / abs(ex.E0),
                ("chiF_rel_err" if mode == "chiF" else "dE0_rel_err"): abs(m["grad"] - want) / abs(want),
                "psi_norm_err": m["psi_norm_err"]}

    main_m = measure(N, k, args.basis, args.steps, args.warmup)
    kernels = kernel_table(main_m)
    sec_per_solve = main_m["ms_total"] / 1e3 / args.steps
    n_loc = main_m["n_loc"]

    dom = kernels.get("reorth_update")
    roofline = None
    if dom and "achieved" in dom:
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written by scripts/ncu_traffic.py from one `ncu --set full` capture
        if os.path.exists(tpath):
            try:
                want = "reorth_update_kernel<float>" if args.basis == "fp32" else "reorth_update_kernel<double>"
                recs = [r for kname, v in json.load(open(tpath)).items() if want in kname for r in v]
                if recs:
                    ratio = sum(r["traffic_over_algorithmic"] for r in recs) / len(recs)
                    traffic = ratio * dom["algorithmic_bytes_per_step"] / dom["launches_per_step"]
                    traffic_src = "profiles/ncu_traffic.json (dram bytes / algorithmic bytes of the captured launches)"
            except Exception:
                traffic = None
        roofline = {"kernel": "reorth_update_kernel (r = u - Q c, pass 2 of full re-orthogonalisation)",
                    "bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "achieved_bytes_per_launch": dom["algorithmic_bytes_per_step"] / dom["launches_per_step"],
                    "bytes_model": "%d n_loc (m + 2) per launch with m stored vectors (read Q[:, :m], read u, write r)"
                                   % (4 if args.basis == "fp32" else 8),
                    "share_of_step": dom["share_of_step"]}

    line = {"metric": METRIC, "value": sec_per_solve, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_m["ms_total"] / args.steps, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": main_m["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": 8,
                    "d2h_bytes_per_step": 16 + 8 * n_loc},
            "gpu_launches": main_m["launches"], "clocks": main_m["clocks"], "roofline": roofline, "kernels": kernels,
            "cg_iterations_per_solve": main_m["cg_iters"], "E0": main_m["E0"],
            ("chiF" if mode == "chiF" else "dE0"): main_m["grad"], "analytic_check": check(main_m)}

    if not args.no_extras:
        if world == 1 and (N, k) != (SAMPLE_N, SAMPLE_K):
            m2 = measure(SAMPLE_N, SAMPLE_K, "fp64", 5, 3, with_e2e=True, with_clocks=False, profile=False)
            line["pairs"] = {"config2": {"workload": workload_name(SAMPLE_N, SAMPLE_K, mode), "per_kernel_events": False,
                                         "ours_s_per_solve": m2["ms_total"] / 1e3 / 5, "ours_e2e_s_per_solve": m2["e2e_s"],
                                         "gpu_launches_per_solve": m2["launches"] / 5, "analytic_check": check(m2)}}
        if args.basis == "fp64" and not args.spins:
            # the same workload with the opt-in fp32 shadow basis (half the reorth traffic, fp64 polish): reported
            # beside the headline, never as the headline (the reference's precision is fp64 throughout)
            m32 = measure(N, k, "fp32", 3, 2, with_e2e=False, with_clocks=False)
            line["fp32_shadow_basis"] = {"config": make_config(N, k, "fp32"), "value": m32["ms_total"] / 1e3 / 3,
                                         "unit": UNIT, "steps": 3, "warmup": 2, "kernels": kernel_table(m32),
                                         "analytic_check": check(m32), "cg_iterations_per_solve": m32["cg_iters"]}
        if world == 1 and mode == "E0_dE0" and not args.spins:
            # SURVEY 8d "work accounting": the E0-only backward has a zero right-hand side; the reference (and the headline
            # above) still runs CG from a random start.  With the opt-in zero start it terminates at once.
            dsea.runtime.cg_start = "zero"
            mz = measure(N, k, args.basis, 3, 1, with_e2e=False, with_clocks=False, profile=False)
            dsea.runtime.cg_start = "random"
            line["zero_start_cg"] = {"value": mz["ms_total"] / 1e3 / 3, "unit": UNIT, "cg_iterations_per_solve": mz["cg_iters"],
                                     "analytic_check": check(mz), "note": "opt-in (DSEA_CG_START=zero); not the headline"}
        if world > 1:
            from dominantsparseeigenad_b200 import selfcheck
            line["selfcheck"] = selfcheck.run()
        head = {}
        if world >= 4 and not args.spins:
            for tag, (Nh, kh, basis) in {"config4_N28_k200": (28, 200, "fp64"),
                                         "config5_N30_k200_fp32basis": (30, 200, "fp32")}.items():
                if Nh == 30 and world < 8:
                    continue
                try:
                    mh = measure(Nh, kh, basis, 2, 1, with_e2e=False, with_clocks=False)
                    head[tag] = {"config": make_config(Nh, kh, basis), "value": mh["ms_total"] / 1e3 / 2, "unit": UNIT,
                                 "steps": 2, "warmup": 1, "kernels": kernel_table(mh), "analytic_check": check(mh),
                                 "cg_iterations_per_solve": mh["cg_iters"]}
                except Exception as exc:                                 # capacity / option not available
                    head[tag] = {"error": str(exc)[:300]}
                    torch.cuda.empty_cache()
        if head:
            line["headline"] = head

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline_leg(args, mode)
        line["cpu_baseline"] = cb
        if "pairs" in line and cb["sample_workload"] == line["pairs"]["config2"]["workload"]:
            line["pairs"]["config2"]["cpu_s_per_solve"] = cb["value"]
            line["pairs"]["config2"]["cpu_kind"] = cb["kind"]
        cached = os.path.join(ROOT, "profiles", "r2_reference_cpu_N24_k200.json")
        if os.path.exists(cached):
            try:
                line["cpu_baseline"]["cached_full_size_run"] = json.load(open(cached))
            except Exception:
                pass
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
