#!/bin/bash
# Run on the GPU box (gpurun): captures the ncu evidence kept under profiles/ (round tag as $1).
R=${1:-r1}
O=gpurun_out
mkdir -p $O
P="ncu --clock-control none"
# launch lists of the timed region only (cudaProfilerStart/Stop in bench.py)
SRJ_CUPROF=1 $P --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_c2_$R.csv python bench.py --rows 20000000 --no-e2e --steps 3 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_c3_from_$R.csv python bench.py --workload c3 --rows 2000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_c3_to_$R.csv python bench.py --workload c3 --direction to_rows --rows 2000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_c4_$R.csv python bench.py --workload c4 --rows 40000000 --no-e2e --steps 3 > /dev/null 2>&1
# one full capture of each top kernel
SRJ_CUPROF=1 $P --profile-from-start off --set full --import-source on -k regex:from_rows_kernel -c 1 -o $O/prof_from_rows_c2_$R python bench.py --rows 20000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --set full --import-source on -k regex:from_rows_kernel -c 1 -o $O/prof_from_rows_c3_$R python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --set full --import-source on -k regex:strings2_kernel -c 1 -o $O/prof_strings2_c3_$R python bench.py --workload c3 --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --set full --import-source on -k regex:from_rows_kernel -c 1 -o $O/prof_from_rows_c4_$R python bench.py --workload c4 --rows 40000000 --no-e2e --steps 1 > /dev/null 2>&1
$P --set full --import-source on -k regex:to_rows2_kernel -s 4 -c 1 -o $O/prof_to_rows2_c2_$R python profiles/time_to_rows.py c2 20000000 > /dev/null 2>&1
SRJ_CUPROF=1 $P --profile-from-start off --set full --import-source on -k regex:to_rows3_kernel -c 1 -o $O/prof_to_rows3_c3_$R python bench.py --workload c3 --direction to_rows --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
SRJ_TR_GENERIC=1 SRJ_CUPROF=1 $P --profile-from-start off --set full --import-source on -k regex:to_rows_kernel -c 1 -o $O/prof_to_rows_generic_c3_$R python bench.py --workload c3 --direction to_rows --rows 1000000 --no-e2e --steps 1 > /dev/null 2>&1
$P --set full --import-source on -k regex:row_hash -s 2 -c 1 -o $O/prof_hash_xx_$R python profiles/time_hash.py 100000000 > /dev/null 2>&1
ls -la $O | tail -20
