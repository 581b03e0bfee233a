#!/bin/bash
# usage: profiles/tune4.sh <rows> "<ENV=VAL ...>" ...   (C3 to_rows, per-kernel times under the ncu launch list)
rows=$1; shift
for cfg in "$@"; do
  env $cfg SRJ_CUPROF=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/l.csv python bench.py --workload c3 --direction to_rows --rows $rows --steps 1 --no-e2e > /tmp/b.log 2>&1 || { echo "cfg $cfg FAILED"; tail -5 /tmp/b.log; continue; }
  python - "$cfg" <<'PY'
import csv,sys
from collections import defaultdict
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10 and r[0].isdigit()]
agg=defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split('(')[0].replace('void ','')[:24]; agg[k][0]+=1; agg[k][1]+=float(r[-1])
print("cfg",sys.argv[1], " | ".join(f"{k} {v[1]/v[0]/1e3:.0f}us" for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:4]))
PY
done
