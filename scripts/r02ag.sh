#!/bin/bash
# generic Stockham transform without run-time divisions: c1 / readme bench + the full GPU suite (generic 1-D and N-D kernels use it)
OUT=gpurun_out; mkdir -p $OUT
for w in c1 readme; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cufft --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$w', '%.4g'%d['value'], d['ms_per_step'])"
done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
