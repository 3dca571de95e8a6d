#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02b_tests.txt
scripts/r02_run.sh r02b "c3 c4" skip
EXB_EPI_BATCH_FASTEST=0 scripts/r02_run.sh r02b_bf0 "c4" skip
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_default.json 2> gpurun_out/r02b_default.err
tail -c 3000 gpurun_out/r02b_default.json; tail -5 gpurun_out/r02b_default.err
