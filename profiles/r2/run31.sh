#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests/test_gpu_unsafe_row.py tests/test_gpu_partition.py -x -q 2>&1 | tail -25 ) > $O/r31_tests.log
tail -25 $O/r31_tests.log
timeout 600 python bench.py --workload partition --steps 5 > $O/r31_partition.log 2>&1; tail -1 $O/r31_partition.log | cut -c1-1500
