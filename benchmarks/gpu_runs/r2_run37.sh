# dev run: small-output 1x1 convolution with 8 lanes per pixel (A/B against one thread per pixel)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_37
timeout 900 python -m pytest tests/test_conv_umma_gpu.py tests/test_networks.py tests/test_ops_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20 > ${O}_tests.log; cat ${O}_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-strict --no-callers --no-e2e --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
SGB_SMALL_G=1 timeout 600 python bench.py --lean > ${O}_bench_g1.log 2>&1
for f in ${O}_bench.log ${O}_bench_g1.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'), (d.get('families') or {}).get('conv_fwd_small'), (s.get('families') or {}).get('conv_fwd_small'))
PY
done
