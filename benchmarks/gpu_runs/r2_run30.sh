# dev run: prefetching FIR kernel for fp32 tensors (A/B)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_30
F="fir_f32_c64_256,fir_f16_c32_1024"
SGB_FIR_PF=2 python benchmarks/prof_shapes.py --reps 5 --cases $F > ${O}_fir_pf2.log 2>&1
SGB_FIR_PF=1 python benchmarks/prof_shapes.py --reps 5 --cases $F > ${O}_fir_pf1.log 2>&1
echo PF2; cat ${O}_fir_pf2.log; echo PF1; cat ${O}_fir_pf1.log
SGB_FIR_PF=2 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k upfirdn 2>&1 | tail -2
SGB_FIR_PF=2 timeout 600 python bench.py --lean > ${O}_bench_pf2.log 2>&1
SGB_FIR_PF=1 timeout 600 python bench.py --lean > ${O}_bench_pf1.log 2>&1
for f in ${O}_bench_pf2.log ${O}_bench_pf1.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); print(sys.argv[1], d['value'], d['ms_per_step'])
PY
done
