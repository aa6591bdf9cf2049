# final code on 4 GPUs (weak scaling, both workloads), as the driver launches it
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 16 --warmup 3 --no-strict --no-callers --no-cpu-baseline --no-roofline > gpurun_out/r2_40_bench_n4.log 2>&1
tail -c 1200 gpurun_out/r2_40_bench_n4.log
