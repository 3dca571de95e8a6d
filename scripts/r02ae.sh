#!/bin/bash
# new pair-formulation tests + c2 A/B of three switches on the ETDRK2 instance
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pairs_are_independent or nyquist_and_full" 2>&1 | tail -12
for v in default b4 nodtw nolines; do
  lib=build/libexb_$v.so; [ $v = default ] && lib=exponax_b200/libexb.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cufft --no-also --no-cpu --no-e2e > $OUT/r02ae_$v.json 2> $OUT/r02ae_$v.err
  python -c "
import json;d=json.loads(open('$OUT/r02ae_$v.json').read().strip().splitlines()[-1]);print('$v', '%.4g'%d['value'], d['ms_per_step'])" 2>&1 | tail -1
done
