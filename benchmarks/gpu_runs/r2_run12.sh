set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --no-strict --no-callers --no-cpu-baseline --no-roofline > gpurun_out/r2_bench_n2.log 2>&1
tail -c 1500 gpurun_out/r2_bench_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n2.log 2>&1
tail -c 600 gpurun_out/r2_bench_ref_n2.log
