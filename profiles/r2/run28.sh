#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_wide.py -x -q -k to_rows 2>&1 | tail -5 ) > $O/r28_tests.log
tail -3 $O/r28_tests.log
SRJ_TW_TRACE=1 timeout 600 python bench.py --direction to_rows --no-e2e --steps 1 --warmup 3 --rows 2000000 2>&1 | grep "^TWF" | tail -12
for f in 12 8; do
SRJ_TW_FILL=$f timeout 600 python bench.py --direction to_rows --no-e2e --steps 3 --rows 10000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['frac'], d['roofline']['ms_per_batch'])"
done
