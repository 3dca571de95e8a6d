#!/bin/bash
# A/B: c2 variants (de-phasing, derived twiddles), c4 row16; parity of the c2 variants on the 1-D tests
mkdir -p gpurun_out
for v in main skew100 skew250 splitbar dtw dtwskew; do
  lib=exponax_b200/libexb.so; [ "$v" != main ] && lib=build/libexb_$v.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-cufft > gpurun_out/r02s_c2_$v.json 2> gpurun_out/r02s_c2_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02s_c2_$v.json').read().strip().splitlines()[-1]); print('$v c2: value %.4g ms %.3f' % (d['value'], d['ms_per_step']))" 2>/dev/null || echo "$v c2 FAILED"
done
for v in main row16; do
  lib=exponax_b200/libexb.so; [ "$v" != main ] && lib=build/libexb_$v.so
  EXB_LIB=$lib timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu --no-e2e --no-cufft > gpurun_out/r02s_c4_$v.json 2> gpurun_out/r02s_c4_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02s_c4_$v.json').read().strip().splitlines()[-1]); print('$v c4: value %.4g ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" 2>/dev/null || echo "$v c4 FAILED"
done
EXB_LIB=build/libexb_dtwskew.so timeout 600 python -m pytest tests -x -q -m gpu -k "1d or burgers or c2" 2>&1 | tail -3
