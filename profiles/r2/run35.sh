#!/bin/bash
# 8 GPUs: the shuffle exchange bench (hash partition -> Kudo split -> all_to_all_single over NVSwitch -> assemble)
O=gpurun_out; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --workload shuffle --gpus 8 --steps 3 --warmup 3 2>&1 | tail -1 > $O/bench_shuffle_n8.log; cut -c1-1400 $O/bench_shuffle_n8.log
