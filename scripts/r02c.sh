#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02c_tests.txt
scripts/r02_run.sh r02c "c2 c3 c4 c1 readme" skip
