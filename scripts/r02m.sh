#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c2 c3 c4" skip
