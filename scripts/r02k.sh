#!/bin/bash
# multi-GPU: slab parity inside pytest + c5 bench lines (usage: r02k.sh TAG NGPUS N)
tag=${1:-r02k}; ng=${2:-2}; N=${3:-1024}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
for mode in "" "--exchange ce"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ng --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --workload c5 --N $N --gpus $ng --steps 3 --warmup 2 $mode > gpurun_out/${tag}_c5_N${N}_n${ng}${mode// /}.json 2> gpurun_out/${tag}_c5_N${N}_n${ng}${mode// /}.err
  tail -c 1200 gpurun_out/${tag}_c5_N${N}_n${ng}${mode// /}.json; echo
  tail -3 gpurun_out/${tag}_c5_N${N}_n${ng}${mode// /}.err
done
