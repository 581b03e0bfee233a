#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts of one kernel from an ncu report (needs -lineinfo + --import-source on).

    python profiles/by_line.py gpurun_out/x.ncu-rep [units_per_launch] [top_n] [kernel-name regex]

Prints executed warp instructions per source line (divided by units_per_launch when given, e.g. tasks per launch),
largest first, plus per-file totals."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
kern = ["-k", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kern, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}
fname = "?"
ie = iss = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ie, iss = r.index("Instructions Executed"), r.index("# Samples")
    elif r[0].isdigit() and ie is not None and len(r) > ie and r[ie] not in ("", "-"):
        try:
            agg[(fname, int(r[0]))] = (int(r[ie]), int(r[iss] or 0), r[1].strip())
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
print(f"total warp instructions {tot}  ({tot / units:.1f} per unit)")
byfile = defaultdict(int)
for (f, l), v in agg.items():
    byfile[f] += v[0]
for f, n in sorted(byfile.items(), key=lambda x: -x[1]):
    print(f"  {f:24s} {n / units:10.2f}")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f:18s}:{l:4d} {v[0] / units:9.2f} {100 * v[0] / tot:5.1f}% samp {v[1]:6d}  {v[2][:100]}")
