#!/bin/bash
# 8-GPU batch-sharded scaling lines: strong (BASELINE's total batch divided over the ranks) and weak
mkdir -p gpurun_out
for sc in strong weak; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 8 --steps 5 --warmup 3 --scaling $sc --no-e2e > gpurun_out/r02t_scale8_$sc.json 2> gpurun_out/r02t_scale8_$sc.err
  python - $sc <<'PY'
import json, sys
sc = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02t_scale8_{sc}.json").read().strip().splitlines()[-1])
    print(sc, "c2 value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k: ("%.4g" % v["value"], "%.3f" % v["roofline"]["frac"]) for k, v in (d.get("also") or {}).items() if "value" in v})
except Exception as e:
    print(sc, "FAILED", e)
PY
done
