#!/bin/bash
# A/B of libexb.so variants on c2 (and optionally other workloads): scripts/run_ab_c2.sh "c2 c1" main nosync ...
wl=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=exponax_b200/libexb.so
  [ "$v" != main ] && lib=build/libexb_$v.so
  for w in $wl; do
    EXB_LIB=$lib timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_${v}_$w.json 2> gpurun_out/ab_${v}_$w.err
    python - "$v" "$w" <<'PY'
import json, sys
v, w = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/ab_{v}_{w}.json").read().strip().splitlines()[-1])
    print(f"{v:10s} {w}: value {d['value']:.4g}  ms/step {d['ms_per_step']:.3f}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(v, w, "FAILED", e)
PY
  done
done
