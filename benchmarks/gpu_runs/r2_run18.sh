set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -rf 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -30
timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown gpurun_out/r2_bd_y.json > gpurun_out/r2_bench_y.log 2>&1
python - <<'PY'
import json
for f in ['gpurun_out/r2_bench_y.log']:
  for ln in open(f):
    if ln.startswith('{'):
        d=json.loads(ln); print(f, 'ffhq256', d['value'], d['ms_per_step'], 'f1024', d['secondary']['value'], d['secondary']['ms_per_step'])
PY
