#!/bin/bash
O=gpurun_out; mkdir -p $O
SRJ_CUPROF=1 timeout 600 ncu --clock-control none --profile-from-start off --set full --import-source on -k regex:to_rows3_kernel -c 1 -o $O/r15_prof_to_rows3 python bench.py --direction to_rows --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
ls -la $O/r15_prof_to_rows3.ncu-rep
