#!/bin/bash
O=gpurun_out; mkdir -p $O
for g in p2p nccl; do
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --rows 20000000 --no-e2e --gather $g 2>&1 | tail -4 ) > $O/r10_bench_n2_$g.log
python - <<PY
import json
for l in open("$O/r10_bench_n2_$g.log"):
    if l.startswith('{'):
        j=json.loads(l); print("$g", j['value'], json.dumps(j['allgather'])[:900])
    elif 'bench:' in l or 'Error' in l or 'error' in l: print(l[:300])
PY
done
