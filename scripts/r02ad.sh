#!/bin/bash
# after the 1-D kernel rewrite: sanitizer on the 1-D set, launch list for c2, full GPU suite, smoke, default bench
OUT=gpurun_out; mkdir -p $OUT
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_workload.py 1d > $OUT/r02ad_sanitizer_${tool}_1d.log 2>&1
  echo "$tool 1d rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload OK' $OUT/r02ad_sanitizer_${tool}_1d.log | tr '\n' ' ')"
done
NCU="ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv"
timeout 300 $NCU -c 40 --log-file $OUT/r02ad_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cufft > /dev/null 2>&1
python scripts/ncu_kernels.py $OUT/r02ad_launches_c2.csv > $OUT/r02ad_launches_c2.txt 2>/dev/null; tail -4 $OUT/r02ad_launches_c2.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $OUT/r02ad_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r02ad_default.json 2> $OUT/r02ad_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02ad_default.json").read().strip().splitlines()[-1])
print("c2 %.4g e2e %.4g cufft %.4g cpu %.4g" % (d["value"], d["e2e"]["value"], d["gpu_library_baseline"]["value"], d["cpu_baseline"]["value"]))
for k, v in d["also"].items():
    print(k, "%.4g" % v["value"], "frac %.3f" % v["roofline"]["frac"])
print(d["clocks"])
PY
