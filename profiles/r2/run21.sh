#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_wide.py tests/test_gpu_to_rows_var.py -x -q 2>&1 | tail -15 ) > $O/r21_tests.log
tail -5 $O/r21_tests.log
rm -f $O/r21_bench.log
for cfg in "8 8" "4 8" "8 4" "12 8" "16 8" "4 4" "6 6"; do
  set -- $cfg
  echo "== FILL=$1 STORE=$2" >> $O/r21_bench.log
  SRJ_TW_FILL=$1 SRJ_TW_STORE=$2 timeout 600 python bench.py --direction to_rows --no-e2e --steps 3 --rows 20000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['frac'], d['roofline']['ms_per_batch'])" >> $O/r21_bench.log
done
cat $O/r21_bench.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'to_rows_wide' --launch-skip 4 -c 2 -o $O/r21_trw python bench.py --direction to_rows --no-e2e --steps 1 --warmup 1 --rows 2000000 > $O/r21_ncu2.log 2>&1
ls -la $O/r21_trw.ncu-rep
