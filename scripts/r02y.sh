#!/bin/bash
# c2 A/B: previous library vs working tree; then the GPU parity suite
mkdir -p gpurun_out
for v in prev default; do
  lib=build/libexb_$v.so; [ $v = default ] && lib=exponax_b200/libexb.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cufft --no-also --no-cpu --no-e2e > gpurun_out/r02y_$v.json 2> gpurun_out/r02y_$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/r02y_$v.json').read().strip().splitlines()[-1]);print('$v', '%.4g'%d['value'], d['ms_per_step'])" 2>&1 | tail -1
done
timeout 900 python -m pytest tests -q -m gpu -k "1d or burgers or rollout or fast" 2>&1 | tail -15
