#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_wide.py -x -q -k to_rows 2>&1 | tail -5 ) > $O/r29_tests.log
tail -3 $O/r29_tests.log
for pf in 4 0 2 8; do
echo "PF=$pf"
SRJ_TW_PF=$pf SRJ_TW_TRACE=1 timeout 600 python bench.py --direction to_rows --no-e2e --steps 1 --warmup 3 --rows 2000000 2>&1 | grep "^TWF" | tail -4
SRJ_TW_PF=$pf timeout 600 python bench.py --direction to_rows --no-e2e --steps 3 --rows 10000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['frac'], d['roofline']['ms_per_batch'])"
done
