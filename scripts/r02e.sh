#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "${2:-2d or kolmogorov or nd or fast or full_size or masked or taylor}" 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "${3:-c3 readme}" skip
EXB_INVPRO_PERSISTENT=0 scripts/r02_run.sh ${tag}_np "c3" skip
