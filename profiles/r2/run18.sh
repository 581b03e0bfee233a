#!/bin/bash
# to_rows_wide: parity first, then to_rows bench with / without the wide kernels, launch list, filler sweep
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_wide.py tests/test_gpu_to_rows_var.py -x -q 2>&1 | tail -25 ) > $O/r18_tests.log
tail -25 $O/r18_tests.log
for off in 0 1; do
  echo "== SRJ_TW_OFF=$off" >> $O/r18_bench.log
  SRJ_TW_OFF=$off timeout 600 python bench.py --direction to_rows --no-e2e --steps 5 2>&1 | tail -3 >> $O/r18_bench.log
done
for f in 8 12 16; do
  echo "== SRJ_TW_FILL=$f" >> $O/r18_bench.log
  SRJ_TW_FILL=$f timeout 600 python bench.py --direction to_rows --no-e2e --steps 5 2>&1 | tail -1 >> $O/r18_bench.log
done
cat $O/r18_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r18_launches_to_rows.csv python bench.py --direction to_rows --no-e2e --steps 1 --warmup 1 --rows 2000000 > $O/r18_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r18_launches_to_rows.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:60]; agg.setdefault(k, []).append(float(r[vi].replace(',', '')))
for k, v in agg.items(): print(f"{k:60s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.1f} us total={sum(v)/1e6:8.2f} ms")
PY
