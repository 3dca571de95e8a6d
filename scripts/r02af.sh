#!/bin/bash
# fixed Nyquist test + ncu full capture of the generic 1-D kernel on c1 (KS, N = 200, one trajectory, 500 steps)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nyquist_and_full" 2>&1 | tail -4
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cufft --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('c1', '%.4g'%d['value'], d['ms_per_step'])"
FULL="ncu --clock-control none --set full --import-source on"
timeout 600 $FULL -k regex:"k1d_kernel" -s 3 -c 1 -f -o $OUT/prof_c1_r02af python bench.py --workload c1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c1_r02af.log 2>&1
python scripts/ncu_stalls.py $OUT/prof_c1_r02af.ncu-rep 2>/dev/null | tail -5
