#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_hash.py -x -q 2>&1 | tail -3 ) > $O/r14_hash_tests.log
for v in "" "SRJ_HASH_STAGES=3" "SRJ_HASH_STAGES=2"; do echo "== $v" >> $O/r14_hash_time.log; ( env $v python profiles/time_hash.py 100000000 2>&1 | head -3 ) >> $O/r14_hash_time.log; done
( timeout 600 python bench.py --direction to_rows --rows 10000000 --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('to_rows c3', j['value'], j['roofline']['frac'], j['roofline']['ms_per_batch'])" ) > $O/r14_torows.log
timeout 600 ncu --clock-control none --set full --import-source on -k regex:row_hash_stream -s 12 -c 1 -o $O/r14_prof_hash_hive python profiles/time_hash.py 100000000 > /dev/null 2>&1
cat $O/r14_hash_tests.log $O/r14_hash_time.log $O/r14_torows.log
