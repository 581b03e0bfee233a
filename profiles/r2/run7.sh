#!/bin/bash
O=gpurun_out; mkdir -p $O
SRJ_CUPROF=1 timeout 600 ncu --clock-control none --cache-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r7_launches_c3.csv python bench.py --workload c3 --rows 4000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 ncu --clock-control none --cache-control none --profile-from-start off --set full --import-source on -k regex:wide_group_scan -c 1 -o $O/r7_prof_scan python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
grep -c . $O/r7_launches_c3.csv
