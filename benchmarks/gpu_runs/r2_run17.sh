set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -rf 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -30
