#!/bin/bash
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c3 readme" skip
EXB_NO_FUSE=1 scripts/r02_run.sh ${tag}_nofuse "c3 readme" skip
