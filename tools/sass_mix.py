#!/usr/bin/env python3
"""Opcode mix and top stall lines from `ncu -i X.ncu-rep --page source --csv` (one kernel)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
src, ex, smp = idx["Source"], idx["Instructions Executed"], idx["# Samples"]
ops, tot, samples = collections.Counter(), 0, 0
lines = []
for r in rows[2:]:
    try:
        n, s = int(r[ex]), int(r[smp])
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[src])
    op = m.group(2) if m else "?"
    ops[op] += n
    tot += n
    samples += s
    lines.append((s, n, r[src].strip()))
print("total warp-instructions", tot, "samples", samples)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%")
print("--- top stall lines")
for s, n, t in sorted(lines, reverse=True)[: int(sys.argv[3]) if len(sys.argv) > 3 else 15]:
    print(f"{100 * s / max(samples, 1):5.1f}%  exec={n:9d}  {t[:90]}")
