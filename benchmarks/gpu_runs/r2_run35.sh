# final code on 2 GPUs (weak scaling, both workloads): bench.py as the driver launches it
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_35
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 16 --warmup 3 --no-strict --no-callers --no-cpu-baseline --no-roofline > ${O}_bench_n2.log 2>&1
tail -c 2500 ${O}_bench_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > ${O}_bench_ref_n2.log 2>&1
tail -c 700 ${O}_bench_ref_n2.log
