#!/bin/bash
# Run on the GPU box (gpurun): the measured numbers quoted in DESIGN.md section 5, saved under gpurun_out/.
# usage: profiles/final_runs.sh [round tag]
R=${1:-r1}
O=gpurun_out
mkdir -p $O
{
  echo "== python bench.py (default: C2 from_rows, N=1)"; python bench.py 2>&1 | tail -1
  echo "== python bench.py --impl reference"; python bench.py --impl reference 2>&1 | tail -1
  echo "== c2 to_rows (profiles/time_to_rows.py)"; python profiles/time_to_rows.py c2 100000000 2>&1 | tail -1
  echo "== c4 from_rows + fused xxhash64"; python bench.py --workload c4 --no-e2e 2>&1 | tail -1
  echo "== c4 from_rows, no hash"; SRJ_BENCH_NOHASH=1 python bench.py --workload c4 --no-e2e 2>&1 | tail -1
  echo "== c4 to_rows"; python profiles/time_to_rows.py c4 400000000 2>&1 | tail -1
  echo "== c3 from_rows"; python bench.py --workload c3 --steps 3 2>&1 | tail -1
  echo "== c3 to_rows"; python bench.py --workload c3 --direction to_rows --steps 3 2>&1 | tail -1
  echo "== hash kernels"; python profiles/time_hash.py 100000000 2>&1 | tail -6
} > $O/final_bench_$R.log 2>&1
cat $O/final_bench_$R.log
