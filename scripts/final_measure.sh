#!/bin/bash
# Final round measurements on one GPU (run under gpurun): smoke, bench lines, ncu launch lists.
TAG=${1:-r01_final}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_c2.json 2> $OUT/${TAG}_bench_c2.err
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_c2_reference.json 2> $OUT/${TAG}_bench_c2_reference.err
for w in c3 c4 c1 readme; do
  timeout 600 python bench.py --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
done
timeout 300 python bench.py --workload readme --cuda-graph --no-cpu > $OUT/${TAG}_bench_readme_graph.json 2> /dev/null
timeout 300 python bench.py --spectral-carry --no-cpu --no-e2e > $OUT/${TAG}_bench_c2_spectral_carry.json 2> /dev/null
NCU="ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv"
timeout 300 $NCU -c 40 --log-file $OUT/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $OUT/${TAG}_launches_c3.csv python bench.py --workload c3 --T 10 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $OUT/${TAG}_launches_c4.csv python bench.py --workload c4 --T 2 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
for f in $OUT/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    r = d.get("roofline") or {}
    print(f"{sys.argv[1].split('/')[-1]:45s} value {d.get('value', 0):.4g} ms {d.get('ms_per_step', 0):.3f} frac {r.get('frac', 0) or 0:.3f} e2e {e.get('value', 0) or 0:.4g} cpu {c.get('value', 0) or 0:.4g}")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
