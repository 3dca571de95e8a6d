#!/bin/bash
# c2 A/B: shuffle unpack (default) vs shared-memory unpack vs derived twiddles; parity of the 1-D tests for the candidates
mkdir -p gpurun_out
for v in noshfl default dtw; do
  lib=build/libexb_$v.so; [ $v = default ] && lib=exponax_b200/libexb.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cufft --no-also --no-cpu --no-e2e > gpurun_out/r02x_$v.json 2> gpurun_out/r02x_$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/r02x_$v.json').read().strip().splitlines()[-1]);print('$v', '%.4g'%d['value'], d['ms_per_step'])" 2>&1 | tail -1
done
timeout 900 python -m pytest tests -x -q -m gpu -k "1d or burgers or rollout or fast" 2>&1 | tail -3
EXB_LIB=build/libexb_dtw.so timeout 900 python -m pytest tests -x -q -m gpu -k "1d or burgers or rollout or fast" 2>&1 | tail -3
