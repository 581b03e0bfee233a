#!/usr/bin/env python
"""Launch lists (ncu --metrics gpu__time_duration.sum --csv) -> profiles/<tag>_launch_lists.md.

    python profiles/summarize_launches.py r1 profiles/launches_c2_r1.csv profiles/launches_c3_from_r1.csv ...
"""
import csv
import os
import sys
from collections import defaultdict

tag, files = sys.argv[1], sys.argv[2:]
out = ["# Launch lists of the timed region (ncu --metrics gpu__time_duration.sum --clock-control none, "
       "cudaProfilerStart/Stop around the timed steps)", "",
       "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", ""]
for f in files:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r[4].split("(")[0].replace("void ", "").replace("srj::", "")
        agg[k][0] += 1
        agg[k][1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    out += ["", f"## {os.path.basename(f)}  (total {tot:.1f} us over {len(rows)} launches)", "",
            "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{tag}_launch_lists.md"), "w").write("\n".join(out) + "\n")
