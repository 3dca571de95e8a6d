set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/slab_multi_gpu_check.py --N 32 > gpurun_out/seg_check32.log 2>&1; echo rc=$?
timeout 400 $TR tests/slab_multi_gpu_check.py --N 128 > gpurun_out/seg_check128.log 2>&1; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --workload c5 --N 1024 --steps 5 --warmup 3 --no-cpu > gpurun_out/seg_c5_raw.log 2>&1; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --workload c5 --N 1024 --steps 5 --warmup 3 --no-cpu --no-raw-exchange > gpurun_out/seg_c5_packed.log 2>&1; echo rc=$?
tail -n 3 gpurun_out/seg_check32.log gpurun_out/seg_check128.log gpurun_out/seg_c5_raw.log gpurun_out/seg_c5_packed.log
