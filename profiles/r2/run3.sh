#!/bin/bash
O=gpurun_out; mkdir -p $O
P="ncu --clock-control none"
( timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_rows.py -x -q 2>&1 | tail -5 ) > $O/r3_tests.log
for v in "" "SRJ_W_WARPS=8" "SRJ_W_WARPS=16" "SRJ_W_SLABCAP=800 SRJ_W_STAGES=4" "SRJ_W_SLABCAP=520"; do
  echo "== $v" >> $O/r3_bench.log
  ( env $v timeout 600 python bench.py --workload c3 --rows 10000000 --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['frac'], j['roofline']['ms_per_batch'])" ) >> $O/r3_bench.log
done
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r3_launches_c3.csv python bench.py --workload c3 --rows 2000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:from_rows_wide_kernel -c 1 -o $O/r3_prof_wideA python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:strings_wide_kernel -c 1 -o $O/r3_prof_wideB python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
cat $O/r3_tests.log $O/r3_bench.log; grep -c . $O/r3_launches_c3.csv
