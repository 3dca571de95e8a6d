#!/bin/bash
# c2 A/B: ETDRK2 instance with register hand-over (straight-line / looped stages) vs previous library; 1-D parity
OUT=gpurun_out; mkdir -p $OUT
for v in prev default o2loop; do
  lib=build/libexb_$v.so; [ $v = default ] && lib=exponax_b200/libexb.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cufft --no-also --no-cpu --no-e2e > $OUT/r02ab_$v.json 2> $OUT/r02ab_$v.err
  python -c "
import json;d=json.loads(open('$OUT/r02ab_$v.json').read().strip().splitlines()[-1]);print('$v', '%.4g'%d['value'], d['ms_per_step'])" 2>&1 | tail -1
done
timeout 900 python -m pytest tests -q -m gpu -k "1d or burgers or rollout or fast" 2>&1 | tail -4
EXB_LIB=build/libexb_o2loop.so timeout 900 python -m pytest tests -q -m gpu -k "1d or burgers or rollout or fast" 2>&1 | tail -3
