#!/bin/bash
# usage: profiles/tune2.sh <workload> <rows> "<variant> <stages> <tile_rows>" ...
wl=$1; rows=$2; shift 2
for cfg in "$@"; do
  set -- $cfg
  SRJ_FR_VARIANT=$1 SRJ_FR_STAGES=$2 SRJ_FR_TILE_ROWS=$3 python bench.py --workload $wl --rows $rows --no-e2e --steps 5 2>&1 | tail -1 | \
    python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('cfg $cfg', 'kernel_ms', r['kernel_ms'], 'GB/s', r['achieved'], 'frac', r['frac'])" || echo "cfg $cfg FAILED"
done
