# dev run: stride-2 convolution on the TMA kernel (element-stride-2 tensor map, column-parity planes)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_29
SGB_TMA_FORCE=1 timeout 600 python -m pytest tests/test_conv_umma_gpu.py tests/test_fused_conv_gpu.py -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests_force.log; cat ${O}_tests_force.log
timeout 900 python -m pytest tests/test_conv_umma_gpu.py tests/test_ops_gpu.py tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests.log; cat ${O}_tests.log
C="conv_s2_f32_c64_256,small_s2_8_n4,small_s2_16_n4,small_s2_8_n32,small_s2_16_n32"
SGB_TMA_S2=1 python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $C > ${O}_s2_tma.log 2>&1
SGB_TMA_S2=0 python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $C > ${O}_s2_halo.log 2>&1
echo TMA; cat ${O}_s2_tma.log; echo HALO; cat ${O}_s2_halo.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
SGB_TMA_S2=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --no-roofline > ${O}_bench_s2off.log 2>&1
for f in ${O}_bench.log ${O}_bench_s2off.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'))
PY
done
