#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "beyond or 1d or burgers or wrong_shape or extension" 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c3" skip
EXB_LIB=build/libexb_noprefetch.so scripts/r02_run.sh ${tag}_noprefetch "c3 c4" skip
