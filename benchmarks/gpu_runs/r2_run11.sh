set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20 > gpurun_out/r2_t11.log; cat gpurun_out/r2_t11.log
timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown gpurun_out/r2_bd_v.json > gpurun_out/r2_bench_v.log 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench_v.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('ffhq256', d['value'], d['ms_per_step'], 'f1024', d['secondary']['value'], d['secondary']['ms_per_step'])
PY
