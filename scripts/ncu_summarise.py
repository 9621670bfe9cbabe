"""Turns the round-2 ncu captures (scripts/profile_round2.sh) into the small tracked summaries under profiles/.

    python scripts/ncu_summarise.py            # reads gpurun_out/r2_*.ncu-rep, r2_launches.csv

  profiles/r2_ncu_<capture>.json   per captured launch: duration, DRAM bytes, DRAM / L2 / L1TEX throughput %, shared
                                   wavefronts, occupancy, registers, dominant stall reasons
  profiles/r2_ncu_<capture>_raw.csv the `--page raw --csv` page itself (every metric)
  profiles/r2_launches_summary.json per-kernel launch count / total time / share of the step from the launch list
  profiles/ncu_traffic.json        DRAM traffic over algorithmic bytes of the re-orthogonalisation kernels (read by
                                   bench.py for roofline.traffic)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
N_LOC = 1 << 24

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def num(v, unit):
    try:
        return float(v.replace(",", "")) * SCALE.get(unit, 1.0)
    except ValueError:
        return v


def raw_page(rep):
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout


def summarise_capture(tag):
    rep = os.path.join(SRC, f"r2_prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return None
    text = raw_page(rep)
    open(os.path.join(OUT, f"r2_ncu_{tag}_raw.csv"), "w").write(text)
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").strip()}
        for m, short in KEEP.items():
            if m in col:
                rec[short] = num(r[col[m]], units[col[m]])
        rec["dram_bytes"] = rec.get("dram_read", 0) + rec.get("dram_write", 0)
        rec["duration_us"] = rec.pop("duration")
        rec["dram_GBps"] = rec["dram_bytes"] / rec["duration_us"] / 1e3
        out.append(rec)
    json.dump(out, open(os.path.join(OUT, f"r2_ncu_{tag}.json"), "w"), indent=1)
    return out


def traffic(recs_by_tag):
    out = {}
    for tag, s in (("reorth_fp64", 8), ("reorth_fp32", 4)):
        for rec in recs_by_tag.get(tag) or []:
            name = rec["kernel"]
            rd = rec["dram_read"]
            if "reorth_update_kernel" in name:          # reads r0 (8) + m columns (s), writes r (8)
                m = max(1, round((rd / N_LOC - 8) / s))
                alg = N_LOC * (s * m + 16.0)
            elif "reorth_dots_kernel" in name:          # reads u (8) + m columns (q_i, q_{i-1} among them, read once), writes r0 (8)
                m = max(1, round((rd / N_LOC - 8) / s))
                alg = N_LOC * (s * m + 16.0)
            else:
                continue
            out.setdefault(name, []).append({"m": m, "dram_bytes": rec["dram_bytes"], "algorithmic_bytes": alg,
                                             "traffic_over_algorithmic": rec["dram_bytes"] / alg,
                                             "duration_us": rec["duration_us"], "dram_GBps": rec["dram_GBps"],
                                             "dram_pct_of_peak": rec.get("dram_pct")})
    json.dump(out, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
    return out


def launches():
    path = os.path.join(SRC, "r2_launches.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").strip()
        unit = r[col["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += num(r[col["Metric Value"]], unit)
    total = sum(v[1] for v in agg.values())
    out = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 python bench.py --steps 1 --warmup 1 "
                      "--no-cpu-baseline --no-extras   (the first 2200 launches = one complete solve and the start of the next; cold-cache serialised times: "
                      "compare SHARES)", "total_us": total,
           "kernels": {k: {"launches": v[0], "total_us": v[1], "share": v[1] / total}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, open(os.path.join(OUT, "r2_launches_summary.json"), "w"), indent=1)
    return out


if __name__ == "__main__":
    recs = {tag: summarise_capture(tag) for tag in ("reorth_fp64", "reorth_fp32", "sweeps_L24", "sweeps_L26", "sweeps_adjoint_L24")}
    for tag, r in recs.items():
        for x in r or []:
            print(tag, x["kernel"][:60], "%.1f us" % x["duration_us"], "dram %.0f MB" % (x["dram_bytes"] / 1e6),
                  "dram%%=%.0f l2%%=%.0f l1tex%%=%.0f" % (x.get("dram_pct", 0), x.get("l2_pct", 0), x.get("l1tex_pct", 0)))
    t = traffic(recs)
    for k, v in t.items():
        print(k, v[0])
    l = launches()
    if l:
        for k, v in list(l["kernels"].items())[:14]:
            print("%-70s %6d %10.0f us %.3f" % (k[:70], v["launches"], v["total_us"], v["share"]))
