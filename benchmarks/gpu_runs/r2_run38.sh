# 2 GPUs: NCCL channel count for the in-graph gradient all-reduce (A/B), final code
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_38
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 16 --warmup 3 --lean > $2 2>&1; python - $2 <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); print(sys.argv[1], d['value'], d['ms_per_step'])
PY
}
run 29541 ${O}_default.log
NCCL_MIN_NCHANNELS=16 run 29542 ${O}_ch16.log
NCCL_MIN_NCHANNELS=32 run 29543 ${O}_ch32.log
SGB_ALLREDUCE_IN_GRAPH=0 run 29544 ${O}_eager_ar.log
