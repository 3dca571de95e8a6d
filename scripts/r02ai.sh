#!/bin/bash
# c5 on all GPUs with the relaxed register caps of the long-line column kernels: pipelined step + phase breakdown
OUT=gpurun_out; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
C5_SKIP_SERIAL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 scripts/c5_phases.py ${N5:-2048} 2>$OUT/r02ai.err | tail -1 > $OUT/r02ai.json; cat $OUT/r02ai.json
