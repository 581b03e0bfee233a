#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_wide.py tests/test_gpu_to_rows_var.py -x -q 2>&1 | tail -15 ) > $O/r22_tests.log
tail -5 $O/r22_tests.log
rm -f $O/r22_bench.log
for cfg in "8 8" "4 8" "8 16" "4 16" "4 4" "8 4"; do
  set -- $cfg
  echo "== FILL=$1 STORE=$2" >> $O/r22_bench.log
  SRJ_TW_FILL=$1 SRJ_TW_STORE=$2 timeout 600 python bench.py --direction to_rows --no-e2e --steps 3 --rows 20000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['frac'], d['roofline']['ms_per_batch'])" >> $O/r22_bench.log
done
cat $O/r22_bench.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'to_rows_wide_fixed' --launch-skip 2 -c 1 -o $O/r22_trw python bench.py --direction to_rows --no-e2e --steps 1 --warmup 1 --rows 2000000 > $O/r22_ncu2.log 2>&1
ls -la $O/r22_trw.ncu-rep
