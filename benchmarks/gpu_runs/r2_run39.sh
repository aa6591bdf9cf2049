# final sanity of the last commits: AugmentPipe parity test, whole GPU suite, smoke, headline bench
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_39
timeout 600 python -m pytest tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf -k augment 2>&1 | grep -E "^E  |^FAILED|passed|failed|Error" | head -20 > ${O}_aug.log; cat ${O}_aug.log
timeout 900 python -m pytest tests -m gpu -q --tb=line -rf 2>&1 | tail -4 > ${O}_tests.log; cat ${O}_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --lean > ${O}_bench.log 2>&1
python - ${O}_bench.log <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); print(sys.argv[1], d['value'], d['ms_per_step'])
PY
