#!/bin/bash
# round 2, GPU call 1: parity of the wide from_rows path + first timings (development build with knobs)
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r1_gpu.txt
( timeout 900 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -25 ) > $O/r1_wide_p2.log
( SRJ_W_FINALIZE=1 timeout 900 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -25 ) > $O/r1_wide_fin.log
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/r1_all.log
for v in "" "SRJ_W_WARPS=11" "SRJ_W_STAGES=2" "SRJ_W_FINALIZE=1" "SRJ_W_SLABCAP=800 SRJ_W_STAGES=4" "SRJ_W_MINROW=100000"; do
  echo "== $v" >> $O/r1_bench.log
  ( env $v timeout 600 python bench.py --workload c3 --rows 10000000 --no-e2e --steps 3 2>&1 | tail -3 ) >> $O/r1_bench.log
done
SRJ_CUPROF=1 timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r1_launches_c3.csv python bench.py --workload c3 --rows 2000000 --no-e2e --steps 1 > /dev/null 2>&1
tail -5 $O/r1_wide_p2.log $O/r1_wide_fin.log $O/r1_all.log; cat $O/r1_bench.log | cut -c1-600
