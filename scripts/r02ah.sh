#!/bin/bash
# c5 on all GPUs of the box: per-phase breakdown (default NCCL settings), then the same with more NCCL p2p channels
OUT=gpurun_out; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $1 scripts/c5_phases.py ${N5:-2048} 2>$OUT/r02ah_$2.err | tail -1 > $OUT/r02ah_$2.json; cat $OUT/r02ah_$2.json; }
run 29541 default
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 run 29542 p2p32
