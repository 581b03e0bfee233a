#!/bin/bash
# Round-2 evidence run: full GPU test suite, the bench lines of every configuration, ncu launch lists and full captures
# of the C3 from_rows kernels.  Outputs -> gpurun_out/ (copied to profiles/ by hand afterwards).
O=gpurun_out; mkdir -p $O
P="ncu --clock-control none"
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/r2_final_tests.log
L=$O/r2_final_bench.log; : > $L
run() { echo "== bench.py $*" >> $L; ( timeout 900 python bench.py "$@" 2>&1 | tail -1 ) >> $L; }
run --steps 10 --warmup 3
run --direction to_rows --steps 5 --no-e2e
run --workload c2 --steps 10
run --workload c4 --steps 5 --no-e2e
for w in nvbench_fixed nvbench_nostr nvbench_var; do for d in to_rows from_rows; do run --workload $w --direction $d --steps 10; done; done
run --workload nvbench_fixed --rows 4194304 --direction from_rows --steps 10
echo "== profiles/time_hash.py 100000000" >> $L; ( python profiles/time_hash.py 100000000 2>&1 | tail -6 ) >> $L
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_c3_from_r2.csv python bench.py --rows 2000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:from_rows_wide_kernel -c 1 -o $O/prof_from_rows_wide_c3_r2 python bench.py --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:strings_wide_kernel -c 1 -o $O/prof_strings_wide_c3_r2 python bench.py --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 timeout 600 $P --profile-from-start off --set full --import-source on -k regex:wide_group_scan -c 1 -o $O/prof_wide_scan_c3_r2 python bench.py --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
timeout 600 $P --set full --import-source on -k regex:row_hash_stream -s 4 -c 1 -o $O/prof_hash_stream_xx_r2 python profiles/time_hash.py 100000000 > /dev/null 2>&1
tail -3 $O/r2_final_tests.log; cut -c1-400 $L
