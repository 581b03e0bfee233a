#!/bin/bash
# 2 GPUs: the NCCL tests of the shard gather and of the shuffle exchange, the shuffle bench and the C5 default bench
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_shuffle.py tests/test_gpu_shard.py -x -q 2>&1 | tail -4 ) > $O/r34_tests_n2.log; cat $O/r34_tests_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --workload shuffle --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 > $O/bench_shuffle_n2.log; cut -c1-1300 $O/bench_shuffle_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e 2>&1 | tail -1 > $O/bench_c5_n2_final.log; cut -c1-400 $O/bench_c5_n2_final.log
