set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 240 $TR tests/slab_multi_gpu_check.py --N 128 > gpurun_out/peer_check128.log 2>&1; echo rc=$?
timeout 200 $TR bench.py --gpus 2 --workload c5 --N 1024 --steps 5 --warmup 3 --no-cpu > gpurun_out/peer_c5_1024.log 2>&1; echo rc=$?
timeout 200 $TR bench.py --gpus 2 --workload c5 --N 1024 --steps 5 --warmup 3 --no-cpu --no-peer-stores > gpurun_out/nopeer_c5_1024.log 2>&1; echo rc=$?
tail -n 2 gpurun_out/peer_check128.log | cut -c1-1500; tail -n 1 gpurun_out/peer_c5_1024.log | cut -c1-1200; tail -n 1 gpurun_out/nopeer_c5_1024.log | cut -c1-400
