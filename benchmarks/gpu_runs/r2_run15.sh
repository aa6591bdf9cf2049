set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20
FIR="fir_f16_c128_256,fir_f16_c32_1024,fir_f32_c64_256"
SGB_FIR_SEP=1 python benchmarks/prof_shapes.py --reps 5 --cases $FIR
SGB_FIR_SEP=0 python benchmarks/prof_shapes.py --reps 5 --cases $FIR
timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --no-roofline > gpurun_out/r2_bench_w.log 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench_w.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('ffhq256', d['value'], d['ms_per_step'], 'f1024', d['secondary']['value'], d['secondary']['ms_per_step'])
PY
