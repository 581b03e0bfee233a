#!/bin/bash
# usage: profiles/tune3.sh <rows> "<variant> <stages> <tile_rows> <stage_kb>" ...   (C3 from_rows, per-kernel times)
rows=$1; shift
for cfg in "$@"; do
  set -- $cfg
  SRJ_FR_VARIANT=$1 SRJ_FR_STAGES=$2 SRJ_FR_TILE_ROWS=$3 SRJ_FR_STAGE_KB=$4 SRJ_CUPROF=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/l.csv python bench.py --workload c3 --rows $rows --steps 1 > /dev/null 2>&1
  python - "$cfg" <<'PY'
import csv,sys
from collections import defaultdict
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10 and r[0].isdigit()]
agg=defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split('(')[0].replace('void ','')[:24]; agg[k][0]+=1; agg[k][1]+=float(r[-1])
print("cfg",sys.argv[1], " | ".join(f"{k} {v[1]/v[0]/1e3:.0f}us" for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:3]))
PY
done
