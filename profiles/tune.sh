#!/bin/bash
# usage: [ENV=...] profiles/tune.sh <workload> <rows> <variants...>   -- prints kernel ms / GB/s / frac per SRJ_FR_VARIANT
wl=$1; rows=$2; shift 2
for v in "$@"; do
  out=$(SRJ_FR_VARIANT=$v python bench.py --workload $wl --rows $rows --no-e2e --steps 5 2>&1 | tail -1)
  echo "$out" | python -c "import sys,json
try:
    j=json.loads(sys.stdin.read()); r=j['roofline']; print('variant $v', 'kernel_ms', r['kernel_ms'], 'GB/s', r['achieved'], 'frac', r['frac'], 'rows/s %.3g' % j['value'], 'box_copy', r.get('this_box_copy_gbs'))
except Exception as e:
    print('variant $v FAILED:', sys.argv[1][-300:])" "$out"
done
