# dev run: tail kernel rewrite (constants in registers, 4 loads in flight), dnoise reduced by shuffles, FIR prefetch default for fp32
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_31
timeout 900 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_tests.log; cat ${O}_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
python - ${O}_bench.log <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'))
        for k in ['scale_bias_act','fused_epilogue_bwd','bias_act_bwd_fused','upfirdn2d','bias_act_g0','bias_act_g1','mod_bwd']:
            print('  ', k, d['families'].get(k), (s.get('families') or {}).get(k))
PY
