#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -15 ) > $O/r19_tests.log
tail -5 $O/r19_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'to_rows|row_size|batch|scan|Scan' -c 80 --csv --log-file $O/r19_launches_to_rows.csv python bench.py --direction to_rows --no-e2e --steps 1 --warmup 1 --rows 2000000 > $O/r19_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r19_launches_to_rows.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:60]; agg.setdefault(k, []).append(float(r[vi].replace(',', '')))
for k, v in agg.items(): print(f"{k:60s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.1f} us total={sum(v)/1e6:8.2f} ms")
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'to_rows_wide' --launch-skip 4 -c 2 -o $O/r19_trw python bench.py --direction to_rows --no-e2e --steps 1 --warmup 1 --rows 2000000 > $O/r19_ncu2.log 2>&1
ls -la $O/r19_trw.ncu-rep
