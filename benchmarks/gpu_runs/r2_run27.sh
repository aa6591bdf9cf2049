# dev run: rewritten pack_weights kernel (parity + timing), TMA kernel forced onto the small layers (A/B)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_27
timeout 900 python -m pytest tests/test_conv_umma_gpu.py tests/test_fused_conv_gpu.py tests/test_networks.py tests/test_ops_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests.log; cat ${O}_tests.log
S="small_s1_8_n4,small_s1_16_n4,small_s1_32_n4,small_s2_8_n4,small_T1_16_n4,small_T2_8_n4,small_s1_8_n32,small_s1_16_n32,small_s1_32_n32,small_T1_16_n32,fwd_f32_c512_32"
python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $S > ${O}_small.log 2>&1
SGB_TMA_FORCE=1 python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $S > ${O}_small_tmaforce.log 2>&1
echo DEFAULT; cat ${O}_small.log; echo TMA_FORCE; cat ${O}_small_tmaforce.log
timeout 600 python bench.py --lean > ${O}_bench.log 2>&1
timeout 600 python bench.py --lean --workload f1024 > ${O}_bench_f1024.log 2>&1
for f in ${O}_bench.log ${O}_bench_f1024.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); print(sys.argv[1], d['value'], d['ms_per_step'])
PY
done
