#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -4 ) > $O/r4_tests.log
for v in "" "SRJ_W_SLABCAP=3200 SRJ_W_STAGES=2" "SRJ_W_SLABCAP=1600 SRJ_W_STAGES=2" "SRJ_W_SLABCAP=1600 SRJ_W_STAGES=3" "SRJ_W_STAGES=4 SRJ_W_ROWS=32" "SRJ_W_WARPS=16" "SRJ_SW_STAGE_KB=30"; do
  echo "== $v" >> $O/r4_bench.log
  ( env $v timeout 600 python bench.py --workload c3 --rows 10000000 --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['frac'], j['roofline']['ms_per_batch'])" ) >> $O/r4_bench.log
done
SRJ_CUPROF=1 timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r4_launches_c3.csv python bench.py --workload c3 --rows 2000000 --no-e2e --steps 1 > /dev/null 2>&1
cat $O/r4_tests.log $O/r4_bench.log; grep -c . $O/r4_launches_c3.csv
