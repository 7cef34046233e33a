#!/usr/bin/env python3
"""Summaries of ncu outputs for profiles/ (run here, no GPU needed).

    python tools/ncu_report.py launches gpurun_out/launches.csv            # per-kernel launches / ms / share
    python tools/ncu_report.py full gpurun_out/prof.ncu-rep [--json out]   # one row per distinct kernel
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name.replace("(int)", "").replace("(bool)", ""))
    return re.sub(r"fp::\(anonymous namespace\)::|fp::<unnamed>::|<?unnamed>::|fp::", "", name)


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= max(ki, vi) or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        n, t = agg.get(short(r[ki]), (0, 0.0))
        agg[short(r[ki])] = (n + 1, t + ms)
    total = sum(t for _, t in agg.values())
    print(f"{sum(n for n, _ in agg.values())} launches, {total:.3f} ms serialised (cold-cache: compare SHARES)\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.3f} | {100 * t / total:.1f}% |")


METRICS = [
    ("time us", "gpu__time_duration.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    ("issue %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("dram read MB", "dram__bytes_read.sum", 1.0),
    ("dram write MB", "dram__bytes_write.sum", 1.0),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("sm GHz", "sm__cycles_elapsed.avg.per_second", 1.0),
]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
TO_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def full(path: str, json_out=None) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen, out = collections.OrderedDict(), []
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        rec = {"kernel": name}
        for label, metric, _ in METRICS:
            if metric not in col:
                continue
            try:
                v = float(r[col[metric]].replace(",", ""))
            except ValueError:
                continue
            u = units[col[metric]]
            if label.endswith("MB"):
                v *= TO_MB.get(u, 1.0)
            if label == "time us":
                v *= TO_US.get(u, 1.0)
            if label == "sm GHz" and u in ("Mhz", "MHz"):
                v /= 1e3
            rec[label] = v
        key = (name, int(rec.get("grid", 0)))
        seen.setdefault(key, []).append(rec)
    labels = [m[0] for m in METRICS]
    print("| kernel | n | " + " | ".join(labels) + " |\n|---|---|" + "---|" * len(labels))
    for (name, _), recs in seen.items():
        avg = {k: sum(r.get(k, 0.0) for r in recs) / len(recs) for k in labels}
        out.append({"kernel": name, "captures": len(recs), **avg})
        cells = []
        for k in labels:
            v = avg[k]
            cells.append(f"{v:.0f}" if k in ("grid", "regs") else (f"{v:.2f}" if k == "sm GHz" else f"{v:.1f}"))
        print(f"| `{name}` | {len(recs)} | " + " | ".join(cells) + " |")
    if json_out:
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "launches":
        launches(path)
    else:
        full(path, sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--json" else None)
