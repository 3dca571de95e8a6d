#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c3 c4 readme" skip
bash scripts/gpu_profile_nd.sh $tag > /dev/null 2>&1
ls -la gpurun_out/*${tag}*
