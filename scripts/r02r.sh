#!/bin/bash
# 8-GPU: slab parity (all modes incl. copy-engine exchange) + c5 2048^3 with the copy-engine exchange
tag=${1:-r02r}; ng=8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "all_gpus" 2>&1 | tail -6 | tee gpurun_out/${tag}_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ng --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --workload c5 --N 2048 --gpus $ng --steps 4 --warmup 2 --exchange ce > gpurun_out/${tag}_c5_N2048_n8_ce.json 2> gpurun_out/${tag}_c5_N2048_n8_ce.err
tail -c 1300 gpurun_out/${tag}_c5_N2048_n8_ce.json; echo; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/${tag}_c5_N2048_n8_ce.err | tail -5
