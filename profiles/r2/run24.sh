#!/bin/bash
O=gpurun_out; mkdir -p $O
rm -f $O/r24_bench.log
for skip in 0 1 2 4 8 16 32 15 31 63; do
  echo "== SKIP=$skip" >> $O/r24_bench.log
  SRJ_TW_SKIP=$skip timeout 600 python bench.py --direction to_rows --no-e2e --steps 3 --rows 10000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['frac'], d['roofline']['ms_per_batch'])" >> $O/r24_bench.log
done
cat $O/r24_bench.log
