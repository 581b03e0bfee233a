#!/bin/bash
# full GPU suite + default bench on the release build (no dev knobs)
O=gpurun_out; mkdir -p $O
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/r30_tests.log
tail -6 $O/r30_tests.log
timeout 900 python bench.py > $O/r30_bench_default.log 2>&1; tail -1 $O/r30_bench_default.log | cut -c1-600
timeout 600 python bench.py --workload nvbench_var --no-e2e --steps 3 > $O/r30_nvbench_var.log 2>&1; tail -1 $O/r30_nvbench_var.log | cut -c1-400
