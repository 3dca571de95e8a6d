#!/bin/bash
# full ncu capture of the lean 1-D kernel at the full c2 batch (T shortened to 200 steps)
OUT=gpurun_out; mkdir -p $OUT
FULL="ncu --clock-control none --set full --import-source on"
timeout 600 $FULL -k regex:"k1d_fast" -s 3 -c 1 -f -o $OUT/prof_c2_r02w python bench.py --workload c2 --T 200 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c2_r02w.log 2>&1
python scripts/ncu_summary.py $OUT/prof_c2_r02w.ncu-rep > $OUT/r02w_full_c2.txt 2>/dev/null
python scripts/ncu_dynmix.py $OUT/prof_c2_r02w.ncu-rep k1d_fast >> $OUT/r02w_full_c2.txt 2>/dev/null
cat $OUT/r02w_full_c2.txt
