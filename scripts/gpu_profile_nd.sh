#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_profile_nd.sh <tag>   -- full ncu captures of the c3 / c4 pass kernels
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:"col_fast|row_fast" -s 40 -c 6 -f -o $OUT/prof_c3_$TAG \
    python bench.py --workload c3 --T 10 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_c3_$TAG.log 2>&1
timeout 600 $NCU -k regex:"col_fast|row_fast" -s 20 -c 10 -f -o $OUT/prof_c4_$TAG \
    python bench.py --workload c4 --T 2 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_c4_$TAG.log 2>&1
ls -la $OUT | tail -5
