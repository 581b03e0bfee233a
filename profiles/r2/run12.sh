#!/bin/bash
O=gpurun_out; mkdir -p $O
N=${1:-8}
for g in p2p nccl; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --rows 40000000 --no-e2e --gather $g 2>&1 | tail -4 ) > $O/r12_bench_n${N}_$g.log
python - <<PY
import json
for l in open("$O/r12_bench_n${N}_$g.log"):
    if l.startswith('{'):
        j=json.loads(l); print("$g", j['value'], j['roofline']['frac'], json.dumps(j['allgather'])[300:1200])
    elif 'bench:' in l or 'rror' in l: print(l[:300])
PY
done
