#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_host_path.py -x -q 2>&1 | tail -15 ) > $O/r16_host_tests.log
( timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('c3', j['value'], j['roofline']['frac'], json.dumps(j['e2e']), json.dumps(j['cpu_baseline'])[:200])" ) > $O/r16_bench_c3.log 2>&1
( timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('c2', j['value'], j['roofline']['frac'], json.dumps(j['e2e']), json.dumps(j['cpu_baseline'])[:200])" ) > $O/r16_bench_c2.log 2>&1
tail -12 $O/r16_host_tests.log; cat $O/r16_bench_c3.log $O/r16_bench_c2.log
