#!/bin/bash
# round 2, first GPU call: baseline lines for c2/c3/c4 on this box + A/B of the 16-points-per-thread row pass (c4, N = 256)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_first_smi.txt
scripts/run_ab_c2.sh "c3 c4" main 2>&1 | tee gpurun_out/r02_first_ab.txt
scripts/run_ab_c2.sh "c4" row16 2>&1 | tee -a gpurun_out/r02_first_ab.txt
EXB_LIB=build/libexb_row16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c4 or navier or 3d or nd" 2>&1 | tail -5 | tee gpurun_out/r02_first_row16_tests.txt
