#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c3 c4" skip
EXB_NO_MASKED_STREAM=1 scripts/r02_run.sh ${tag}_noms "c3 c4" skip
bash scripts/run_sanitizer.sh $tag
