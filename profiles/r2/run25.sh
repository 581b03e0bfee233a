#!/bin/bash
O=gpurun_out; mkdir -p $O
rm -f $O/r25_bench.log
( timeout 1200 python -m pytest tests/test_gpu_wide.py -x -q -k to_rows 2>&1 | tail -5 ) > $O/r25_tests.log
tail -3 $O/r25_tests.log
for cfg in "3200 8" "1600 8" "1600 3" "800 8" "800 4" "1200 8"; do
  set -- $cfg
  echo "== SLABCAP=$1 STAGES=$2" >> $O/r25_bench.log
  SRJ_TW_SLABCAP=$1 SRJ_TW_STAGES=$2 timeout 600 python bench.py --direction to_rows --no-e2e --steps 3 --rows 10000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['frac'], d['roofline']['ms_per_batch'])" >> $O/r25_bench.log
done
cat $O/r25_bench.log
