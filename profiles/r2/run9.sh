#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -8 ) > $O/r9_shard_tests.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --rows 20000000 --no-e2e 2>&1 | tail -3 ) > $O/r9_bench_n2.log
( timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 ) > $O/r9_bench_n1.log
( timeout 900 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 ) > $O/r9_bench_ref.log
tail -3 $O/r9_shard_tests.log; cut -c1-1500 $O/r9_bench_n2.log; cut -c1-3000 $O/r9_bench_n1.log; cut -c1-800 $O/r9_bench_ref.log
