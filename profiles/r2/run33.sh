#!/bin/bash
O=gpurun_out; mkdir -p $O
SRJ_TH_STR_ROWS=20000000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'row_hash_kernel' -c 1 -o $O/r33_hash_str python profiles/time_hash.py 20000000 > $O/r33.log 2>&1
ls -la $O/r33_hash_str.ncu-rep
