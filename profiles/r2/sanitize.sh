#!/bin/bash
# compute-sanitizer over a reduced set of the GPU parity tests (memcheck, then racecheck); logs -> profiles/r2_sanitizer.log
O=gpurun_out; mkdir -p $O
SEL_WIDE='(c3-33 or c3-129 or odd-65 or many_strings-64 or one_slab-127) and test_wide_from_rows'
{
echo "### compute-sanitizer --tool memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file $O/memcheck_raw.log python -m pytest -q -x tests/test_gpu_wide.py -k "$SEL_WIDE or (unaligned and c3 and (1 or 8)) or non_canonical or long_strings or exact_size" tests/test_gpu_rows.py -k "$SEL_WIDE or (unaligned and c3) or non_canonical or long_strings or exact_size or (fixed_width and c2-257) or (strings_both and (mixed-33 or c3_small-1000)) or fused" 2>&1 | tail -3
echo "memcheck summary:"; grep -E "ERROR SUMMARY|Invalid|misaligned|out of bounds" $O/memcheck_raw.log | sort | uniq -c | head -20
echo "### compute-sanitizer --tool racecheck"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 --log-file $O/racecheck_raw.log python -m pytest -q -x tests/test_gpu_wide.py tests/test_gpu_rows.py -k "(c3-33 and test_wide_from_rows) or (odd-65 and test_wide_from_rows) or (fixed_width and c2-257) or (strings_both and mixed-33)" 2>&1 | tail -3
echo "racecheck summary:"; grep -E "RACECHECK SUMMARY|hazard" $O/racecheck_raw.log | sort | uniq -c | head -20
echo "### compute-sanitizer --tool synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 --log-file $O/synccheck_raw.log python -m pytest -q -x tests/test_gpu_wide.py -k "c3-129 and test_wide_from_rows" 2>&1 | tail -3
grep -E "ERROR SUMMARY" $O/synccheck_raw.log
} > $O/r2_sanitizer.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_more.py -q -k "2gib" 2>&1 | tail -3 ) >> $O/r2_sanitizer.log
cat $O/r2_sanitizer.log
