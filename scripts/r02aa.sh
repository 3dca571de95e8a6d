#!/bin/bash
# c2 A/B: phase-wise multi-line transforms vs previous library; 1-D parity; ncu full capture of the new default
OUT=gpurun_out; mkdir -p $OUT
for v in prev default; do
  lib=build/libexb_$v.so; [ $v = default ] && lib=exponax_b200/libexb.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cufft --no-also --no-cpu --no-e2e > $OUT/r02aa_$v.json 2> $OUT/r02aa_$v.err
  python -c "
import json;d=json.loads(open('$OUT/r02aa_$v.json').read().strip().splitlines()[-1]);print('$v', '%.4g'%d['value'], d['ms_per_step'])" 2>&1 | tail -1
done
timeout 900 python -m pytest tests -q -m gpu -k "1d or burgers or rollout or fast" 2>&1 | tail -4
FULL="ncu --clock-control none --set full --import-source on"
timeout 600 $FULL -k regex:"k1d_fast" -s 3 -c 1 -f -o $OUT/prof_c2_r02aa python bench.py --workload c2 --T 200 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c2_r02aa.log 2>&1
python scripts/ncu_stalls.py $OUT/prof_c2_r02aa.ncu-rep 2>/dev/null | tail -5
