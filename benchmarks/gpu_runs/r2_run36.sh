# verification of the last kernel change (tile decode in publish only when a pass runs) + part B of the record: ncu launch list, main phases
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_36
timeout 900 python -m pytest tests/test_conv_umma_gpu.py tests/test_fused_conv_gpu.py tests/test_networks.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20 > ${O}_tests.log; cat ${O}_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-strict --no-callers --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
python - ${O}_bench.log <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], 'f1024', s.get('value'), s.get('ms_per_step'))
PY
bash benchmarks/gpu_runs/r2_run34.sh
