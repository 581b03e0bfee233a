#!/bin/bash
O=gpurun_out; mkdir -p $O
run() { echo "== $*" >> $O/r5_bench.log; ( env "$@" timeout 600 python bench.py --workload c3 --rows 10000000 --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['frac'], j['roofline']['ms_per_batch'], j.get('phases'))" ) >> $O/r5_bench.log; }
run SRJ_BENCH_PHASES=1
run SRJ_BENCH_OVERLAP=1
A1="SRJ_W_SLABCAP=3200 SRJ_W_STAGES=2"
for g in 48 56 64 72 80; do
  run SRJ_BENCH_OVERLAP=1 SRJ_W_GRID=$g SRJ_SW_GRID=$((148-g))
  run SRJ_BENCH_OVERLAP=1 SRJ_W_GRID=$g SRJ_SW_GRID=$((148-g)) $A1
done
cat $O/r5_bench.log
