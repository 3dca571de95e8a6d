#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "3d or navier or c4 or masked or nd or slab or fast" 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
scripts/r02_run.sh $tag "c4" skip
EXB_PLAIN_TMA=0 scripts/r02_run.sh ${tag}_notma "c4" skip
bash scripts/gpu_profile_nd.sh $tag > /dev/null 2>&1
ls -la gpurun_out/*${tag}* | head
