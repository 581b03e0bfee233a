#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_hash.py -x -q 2>&1 | tail -6 ) > $O/r13_hash_tests.log
( python profiles/time_hash.py 100000000 2>&1 | tail -8 ) > $O/r13_hash_time.log
( SRJ_HASH_NOSTREAM=1 python profiles/time_hash.py 100000000 2>&1 | head -3 ) > $O/r13_hash_time_old.log
timeout 600 ncu --clock-control none --set full --import-source on -k regex:row_hash_stream -s 4 -c 1 -o $O/r13_prof_hash python profiles/time_hash.py 100000000 > /dev/null 2>&1
cat $O/r13_hash_tests.log $O/r13_hash_time.log $O/r13_hash_time_old.log
