# dev run: one-pass kernels sized to one resident wave (occupancy query)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_32
timeout 900 python -m pytest tests/test_fused_conv_gpu.py tests/test_networks.py tests/test_ops_gpu.py tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests.log; cat ${O}_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
python - ${O}_bench.log <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'))
        for k in ['scale_bias_act','fused_epilogue_bwd','bias_act_bwd_fused','mod_bwd']:
            a=d['families'].get(k) or {}; b=(s.get('families') or {}).get(k) or {}
            print('  ', k, round(a.get('ms_per_step',0),3), round(a.get('frac_hbm',0),3), '|', round(b.get('ms_per_step',0),3), round(b.get('frac_hbm',0),3))
PY
