#!/bin/bash
O=gpurun_out; mkdir -p $O
SRJ_TW_TRACE=1 timeout 600 python bench.py --direction to_rows --no-e2e --steps 1 --warmup 3 --rows 2000000 2>&1 | grep "^TWF" | tail -400 > $O/r27_trace.log
wc -l $O/r27_trace.log
