#!/bin/bash
O=gpurun_out; mkdir -p $O
P="ncu --clock-control none"
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:from_rows_wide_kernel -c 1 -o $O/r2_prof_wideA python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:strings_wide_kernel -c 1 -o $O/r2_prof_wideB python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
for v in "" "SRJ_W_WARPS=7" "SRJ_W_SLABCAP=520" "SRJ_W_SLABCAP=1600 SRJ_W_STAGES=2" "SRJ_SW_STAGE_KB=28" "SRJ_W_SLABCAP=520 SRJ_W_WARPS=11"; do
  echo "== $v" >> $O/r2_bench.log
  ( env $v timeout 600 python bench.py --workload c3 --rows 10000000 --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['frac'], j['roofline']['ms_per_batch'])" ) >> $O/r2_bench.log
done
cat $O/r2_bench.log
