#!/bin/bash
# Round-2 final evidence run (after the f1 / f3 work): full GPU suite, every bench line, launch lists and ncu captures of
# the new kernels, compute-sanitizer over the new paths.  Outputs -> gpurun_out/ (copied to profiles/ afterwards).
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
run --workload partition --steps 5
run --workload unsafe_c2 --steps 5
run --workload unsafe_c2 --direction to_rows --steps 5
run --workload kudo --direction to_rows --steps 5
run --workload kudo --steps 5
run --workload shuffle --steps 3
echo "== profiles/time_hash.py 100000000" >> $L; ( python profiles/time_hash.py 100000000 2>&1 | tail -6 ) >> $L
timeout 600 $P --metrics gpu__time_duration.sum -k regex:'part_|partition_|i32_|murmur|row_hash|gather_|scatter_' --csv --log-file $O/launches_partition_r2.csv python bench.py --workload partition --rows 20000000 --steps 1 --warmup 3 > /dev/null 2>&1
timeout 600 $P --set full --import-source on -k regex:partition_move_tile -s 1 -c 1 -o $O/prof_partition_move_r2 python bench.py --workload partition --rows 20000000 --steps 1 --warmup 3 > /dev/null 2>&1
timeout 600 $P --set full --import-source on -k regex:ur_from_rows -s 1 -c 1 -o $O/prof_ur_from_rows_r2 python bench.py --workload unsafe_c2 --rows 20000000 --steps 1 --warmup 3 > /dev/null 2>&1
timeout 600 $P --set full --import-source on -k regex:ur_to_rows -s 1 -c 1 -o $O/prof_ur_to_rows_r2 python bench.py --workload unsafe_c2 --direction to_rows --rows 20000000 --steps 1 --warmup 3 > /dev/null 2>&1
timeout 600 $P --set full --import-source on -k regex:kudo_split_kernel -s 1 -c 1 -o $O/prof_kudo_split_r2 python bench.py --workload kudo --direction to_rows --rows 20000000 --steps 1 --warmup 3 > /dev/null 2>&1
timeout 600 $P --set full --import-source on -k regex:kudo_assemble_kernel -s 1 -c 1 -o $O/prof_kudo_assemble_r2 python bench.py --workload kudo --rows 20000000 --steps 1 --warmup 3 > /dev/null 2>&1
{
echo "### compute-sanitizer --tool memcheck (partition.cu, unsafe_row.cu, kudo.cu)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file $O/memcheck2_raw.log python -m pytest -q -x tests/test_gpu_partition.py tests/test_gpu_unsafe_row.py tests/test_gpu_kudo.py -k "(matches_oracle and (4097 or 1000 or 33)) or string_and_mixed or empty or by_id or long_and_empty or decimal128 or without_row_offsets or test_split_bytes or no_masks or round_trip or two_tables" 2>&1 | tail -3
echo "memcheck summary:"; grep -E "ERROR SUMMARY|Invalid|misaligned|out of bounds" $O/memcheck2_raw.log | sort | uniq -c | head -20
echo "### compute-sanitizer --tool racecheck"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 --log-file $O/racecheck2_raw.log python -m pytest -q -x tests/test_gpu_partition.py tests/test_gpu_unsafe_row.py tests/test_gpu_kudo.py -k "(test_hash_partition_matches_oracle and 4097 and (200 or 7)) or (test_unsafe_rows_match_oracle and mixed and 1000) or by_id or (test_split_bytes and 1000) or test_assemble_round_trip" 2>&1 | tail -3
echo "racecheck summary:"; grep -E "RACECHECK SUMMARY|hazard" $O/racecheck2_raw.log | sort | uniq -c | head -20
} > $O/r2_sanitizer2.log 2>&1
tail -3 $O/r2_final_tests.log; cut -c1-260 $L; cat $O/r2_sanitizer2.log
