#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_rows.py -x -q 2>&1 | tail -12 ) > $O/r8_tests.log
run() { echo "== $*" >> $O/r8_bench.log; ( env "$@" timeout 600 python bench.py --workload c3 --rows 10000000 --no-e2e --steps 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['frac'], j['roofline']['ms_per_batch'], j.get('phases'))" ) >> $O/r8_bench.log; }
run SRJ_BENCH_PHASES=1
run SRJ_BENCH_PHASES=1 SRJ_W_SLABCAP=3200 SRJ_W_STAGES=2
run SRJ_BENCH_OVERLAP=1
run SRJ_BENCH_OVERLAP=1 SRJ_W_SLABCAP=3200 SRJ_W_STAGES=2
cat $O/r8_tests.log $O/r8_bench.log
