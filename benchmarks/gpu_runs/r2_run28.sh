# dev run: TMA kernel on 16x16 / 32x32 layers, deeper pipeline in the 16-bit halo weight gradient, prefetching FIR kernel (16-bit)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_28
timeout 900 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests.log; cat ${O}_tests.log
F="fir_f16_c128_256,fir_f16_c32_1024,fir_f32_c64_256"
SGB_FIR_PF=1 python benchmarks/prof_shapes.py --reps 5 --cases $F > ${O}_fir_pf1.log 2>&1
SGB_FIR_PF=0 python benchmarks/prof_shapes.py --reps 5 --cases $F > ${O}_fir_pf0.log 2>&1
echo PF1; cat ${O}_fir_pf1.log; echo PF0; cat ${O}_fir_pf0.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
SGB_FIR_PF=0 timeout 600 python bench.py --lean --workload f1024 > ${O}_bench_f1024_pf0.log 2>&1
SGB_WGRAD_LA=1 timeout 600 python bench.py --lean --workload f1024 > ${O}_bench_f1024_la1.log 2>&1
for f in ${O}_bench.log ${O}_bench_f1024_pf0.log ${O}_bench_f1024_la1.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'))
PY
done
