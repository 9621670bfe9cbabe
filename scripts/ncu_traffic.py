"""Summarises an `ncu --set full` raw CSV page into profiles/ncu_traffic.json.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_traffic.py raw.csv --n 16777216 --out profiles/ncu_traffic.json

For every captured launch it records DRAM bytes (read + write), duration and, for the two reorth kernels,
the number of stored vectors m inferred from the read volume (reads = 8 n (m + 1 [+3 recurrence vectors])),
the algorithmic bytes of THAT launch and the ratio traffic / algorithmic.  bench.py multiplies its average
algorithmic bytes per launch by that ratio to fill `roofline.traffic`.
"""
import argparse
import csv
import json


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--n", type=int, required=True, help="rows per rank (n_loc) of the profiled run")
    ap.add_argument("--out", default="profiles/ncu_traffic.json")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        rec = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr, "duration_us": us,
               "dram_GBps": (rd + wr) / us / 1e3,
               "dram_pct_of_peak": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])}
        if "reorth_update_kernel" in name:          # reads u/r0 + m columns, writes r
            m = max(1, round(rd / (8 * a.n)) - 1)
            rec.update(m=m, algorithmic_bytes=8.0 * a.n * (m + 2))
        elif "reorth_dots_kernel" in name:          # reads u, q_i, q_{i-1} + m columns, writes r0
            m = max(1, round(rd / (8 * a.n)) - 3)
            rec.update(m=m, algorithmic_bytes=8.0 * a.n * (m + 4))
        elif "tfim_sweep" in name:
            rec.update(algorithmic_bytes=16.0 * a.n, note="algorithmic bytes are per MATVEC (all sweeps together)")
        if "algorithmic_bytes" in rec:
            rec["traffic_over_algorithmic"] = rec["dram_bytes"] / rec["algorithmic_bytes"]
        out.setdefault(name, []).append(rec)
    json.dump(out, open(a.out, "w"), indent=1)
    for k, v in out.items():
        print(k, json.dumps(v[0]))


if __name__ == "__main__":
    main()
