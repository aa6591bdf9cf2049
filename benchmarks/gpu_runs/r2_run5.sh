set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|^tests/.*Error|passed|failed|config A|\(worst" | head -80 > gpurun_out/r2_t5.log
cat gpurun_out/r2_t5.log
