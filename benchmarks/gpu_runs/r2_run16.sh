set -x
mkdir -p gpurun_out
SGB_TMA_FORCE=1 timeout 900 python -m pytest tests/test_fused_conv_gpu.py tests/test_conv_umma_gpu.py -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20
timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown gpurun_out/r2_bd_x.json > gpurun_out/r2_bench_x.log 2>&1
SGB_FUSED_CONV=0 timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --no-roofline > gpurun_out/r2_bench_x0.log 2>&1
python - <<'PY'
import json
for f in ['gpurun_out/r2_bench_x.log','gpurun_out/r2_bench_x0.log']:
  for ln in open(f):
    if ln.startswith('{'):
        d=json.loads(ln); print(f, 'ffhq256', d['value'], d['ms_per_step'], 'f1024', d['secondary']['value'], d['secondary']['ms_per_step'])
PY
