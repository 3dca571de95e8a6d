#!/bin/bash
# Round-2 final measurements on one GPU (run under gpurun): smoke, bench lines, ncu launch lists + full captures.
TAG=${1:-r02_final}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
for w in c1 readme; do
  timeout 600 python bench.py --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
done
timeout 300 python bench.py --workload readme --cuda-graph --no-cpu --no-cufft > $OUT/${TAG}_bench_readme_graph.json 2> /dev/null
timeout 300 python bench.py --workload c2 --spectral-carry --no-cpu --no-e2e --no-cufft > $OUT/${TAG}_bench_c2_spectral_carry.json 2> /dev/null
for v in nosync nounify; do
  EXB_LIB=build/libexb_$v.so timeout 300 python bench.py --workload c2 --no-cpu --no-e2e --no-cufft > $OUT/${TAG}_bench_c2_$v.json 2> /dev/null
done
NCU="ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv"
timeout 300 $NCU -c 40 --log-file $OUT/${TAG}_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-cufft > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $OUT/${TAG}_launches_c3.csv python bench.py --workload c3 --T 10 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > /dev/null 2>&1
timeout 300 $NCU -c 400 --log-file $OUT/${TAG}_launches_c4.csv python bench.py --workload c4 --T 2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > /dev/null 2>&1
FULL="ncu --clock-control none --set full --import-source on"
timeout 600 $FULL -k regex:"k1d_fast" -s 3 -c 1 -f -o $OUT/prof_c2_$TAG python bench.py --workload c2 --batch 2048 --T 200 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c2_$TAG.log 2>&1
timeout 600 $FULL -k regex:"col_|row_fast|etdrk_masked" -s 30 -c 8 -f -o $OUT/prof_c3_$TAG python bench.py --workload c3 --T 10 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c3_$TAG.log 2>&1
timeout 600 $FULL -k regex:"col_|row_fast|etdrk_masked" -s 26 -c 14 -f -o $OUT/prof_c4_$TAG python bench.py --workload c4 --T 2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-cufft > $OUT/ncu_c4_$TAG.log 2>&1
# text summaries of the full captures on the box (the .ncu-rep files together exceed gpurun's 64 MiB return limit)
for w in c2 c3 c4; do
  python scripts/ncu_stalls.py $OUT/prof_${w}_$TAG.ncu-rep > $OUT/${TAG}_full_$w.txt 2>/dev/null
  python scripts/ncu_dynmix.py $OUT/prof_${w}_$TAG.ncu-rep >> $OUT/${TAG}_full_$w.txt 2>/dev/null
done
python scripts/ncu_lines.py $OUT/prof_c3_$TAG.ncu-rep "row_fast_kernel" 25 --wf > $OUT/${TAG}_lines_c3_row.txt 2>/dev/null
python scripts/ncu_lines.py $OUT/prof_c4_$TAG.ncu-rep "col_plain_tma_kernel" 25 > $OUT/${TAG}_lines_c4_tma.txt 2>/dev/null
python scripts/ncu_lines.py $OUT/prof_c2_$TAG.ncu-rep "k1d_fast_kernel" 30 > $OUT/${TAG}_lines_c2.txt 2>/dev/null
ncu -i $OUT/prof_c2_$TAG.ncu-rep --page details 2>/dev/null | head -150 > $OUT/${TAG}_details_c2.txt
rm -f $OUT/prof_c3_$TAG.ncu-rep $OUT/prof_c4_$TAG.ncu-rep
for w in c2 c3 c4; do python scripts/ncu_kernels.py $OUT/${TAG}_launches_$w.csv > $OUT/${TAG}_launches_$w.txt 2>/dev/null; done
for f in $OUT/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    r = d.get("roofline") or {}
    g = d.get("gpu_library_baseline") or {}
    print(f"{sys.argv[1].split('/')[-1]:45s} value {d.get('value', 0):.4g} ms {d.get('ms_per_step', 0):.3f} frac {r.get('frac', 0) or 0:.3f} e2e {e.get('value', 0) or 0:.4g} cpu {c.get('value', 0) or 0:.4g} cufft {g.get('value', 0) or 0:.4g}")
    for k, v in (d.get("also") or {}).items():
        print(f"    also {k}: value {v.get('value', 0):.4g} ms {v.get('ms_per_step', 0):.3f} frac {(v.get('roofline') or {}).get('frac', 0):.3f} cufft {((v.get('gpu_library_baseline') or {}).get('value') or 0):.4g}")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
ls -la $OUT/*${TAG}* | head -40
