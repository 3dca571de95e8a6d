#!/bin/bash
# round 2 GPU call: full GPU parity suite, then c3 / c4 (/ more) bench lines of the default library
# usage: scripts/r02_run.sh TAG "c3 c4" [pytest -k expr]
tag=$1; wl=${2:-"c3 c4"}; kexpr=$3
mkdir -p gpurun_out
if [ "$kexpr" != "skip" ]; then
  if [ -n "$kexpr" ]; then
    timeout 1500 python -m pytest tests -x -q -m gpu -k "$kexpr" 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.txt
  else
    timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.txt
  fi
fi
python scripts/peaks.py 2>&1 | tee gpurun_out/${tag}_peaks.json
for w in $wl; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
  python - "$tag" "$w" <<'PY'
import json, sys
v, w = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/{v}_{w}.json").read().strip().splitlines()[-1])
    print(f"{v:10s} {w}: value {d['value']:.4g}  ms/step {d['ms_per_step']:.3f}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(v, w, "FAILED", e)
PY
done
