#!/bin/bash
# ncu full capture of the ETDRK2 1-D instance at the full c2 batch (T = 200)
OUT=gpurun_out; mkdir -p $OUT
FULL="ncu --clock-control none --set full --import-source on"
timeout 600 $FULL -k regex:"k1d_fast" -s 3 -c 1 -f -o $OUT/prof_c2_r02ac python bench.py --workload c2 --T 200 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c2_r02ac.log 2>&1
python scripts/ncu_stalls.py $OUT/prof_c2_r02ac.ncu-rep 2>/dev/null | tail -5
