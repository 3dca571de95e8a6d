#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "beyond or wave or 1d or burgers or extension" 2>&1 | tail -12 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c3 c4" skip
EXB_LIB=build/libexb_noprefetch.so scripts/r02_run.sh ${tag}_noprefetch "c3 c4" skip
