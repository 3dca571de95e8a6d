#!/bin/bash
# confirm run: full GPU suite + the driver's default bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02u_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02u_default.json 2> gpurun_out/r02u_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02u_default.json").read().strip().splitlines()[-1])
print("c2 %.4g e2e %.4g cufft %.4g cpu %.4g" % (d["value"], d["e2e"]["value"], d["gpu_library_baseline"]["value"], d["cpu_baseline"]["value"]))
for k, v in d["also"].items():
    print(k, "%.4g" % v["value"], "frac %.3f" % v["roofline"]["frac"], "cufft %.4g" % v["gpu_library_baseline"]["value"])
print(d["roofline"]["bound"]); print(d["clocks"])
PY
