# dev run: per-tile (instead of per-K-block) tile decode in the halo kernel's producers; incremental cursors in the 16-bit wgrad kernel
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_25
timeout 900 python -m pytest tests/test_conv_umma_gpu.py tests/test_fused_conv_gpu.py tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests.log; cat ${O}_tests.log
S4="small_s1_8_n4,small_s1_16_n4,small_s1_32_n4,small_s2_8_n4,small_s2_16_n4,small_T1_16_n4,small_T2_8_n4"
S32="small_s1_8_n32,small_s1_16_n32,small_s1_32_n32,small_s2_8_n32,small_s2_16_n32,small_T1_16_n32,small_T2_8_n32"
python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $S4,$S32 > ${O}_small_graph.log 2>&1
CASES="fwd_f32_c64_256_n32,fwd_f32_c512_32,fwd_f16_c32_1024,fwd_f16_c64_512,wgrad_f32_c64_256,wgrad_f32_c512_32,wgrad_f16_c32_1024,wgrad_f16_c64_512,convT_s2_f32_c128_128,conv_s2_f32_c64_256"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > ${O}_prof_plain.log 2>&1
cat ${O}_small_graph.log ${O}_prof_plain.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
python - ${O}_bench.log <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'))
PY
find gpurun_out -type f -size +8M -print -delete
