#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_profile.sh <tag>
# Writes launch lists (gpu__time_duration) and one `--set full` capture per dominant kernel into gpurun_out/.
set -x
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: compare shares, not absolutes)
$NCU --metrics gpu__time_duration.sum -c 40 --csv --log-file $OUT/launches_c2_$TAG.csv \
    python bench.py --workload c2 --batch 16384 --T 100 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_c2_$TAG.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $OUT/launches_c3_$TAG.csv \
    python bench.py --workload c3 --batch 128 --T 2 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_c3_$TAG.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $OUT/launches_c4_$TAG.csv \
    python bench.py --workload c4 --batch 4 --T 1 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_c4_$TAG.log 2>&1
# full captures of the dominant kernels
$NCU --set full --import-source on -k regex:k1d -s 3 -c 1 -f -o $OUT/prof_c2_$TAG \
    python bench.py --workload c2 --batch 16384 --T 20 --steps 1 --warmup 3 --no-e2e --no-cpu >> $OUT/ncu_c2_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:"col_pass|row_pass" -s 60 -c 6 -f -o $OUT/prof_c3_$TAG \
    python bench.py --workload c3 --batch 128 --T 2 --steps 1 --warmup 3 --no-e2e --no-cpu >> $OUT/ncu_c3_$TAG.log 2>&1
ls -la $OUT
