#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'partition_move' --launch-skip 1 -c 1 -o $O/r32_move python bench.py --workload partition --steps 1 --warmup 3 --rows 20000000 > $O/r32_ncu.log 2>&1
ls -la $O/r32_move.ncu-rep
