#!/bin/bash
# compute-sanitizer over the fast kernels (run under gpurun): racecheck + synccheck (+ memcheck on the 2-D / 1-D set)
tag=${1:-r02}
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  for part in 1d 2d 3d; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_workload.py $part > gpurun_out/${tag}_sanitizer_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload OK' gpurun_out/${tag}_sanitizer_${tool}_${part}.log | tr '\n' ' ')"
  done
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_workload.py 2d > gpurun_out/${tag}_sanitizer_memcheck_2d.log 2>&1
echo "memcheck 2d rc=$? $(grep -E 'ERROR SUMMARY|sanitize workload OK' gpurun_out/${tag}_sanitizer_memcheck_2d.log | tr '\n' ' ')"
