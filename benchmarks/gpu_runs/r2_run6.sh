set -x
mkdir -p gpurun_out
SGB_TMA_BO=1 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_tma_bo1.log 2>&1
SGB_TMA_BO=0 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_tma_bo0.log 2>&1
SGB_TMA=0 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_tma_off.log 2>&1
timeout 1500 python -m pytest tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|passed|failed|config A|\(worst" | head -60 > gpurun_out/r2_t6.log
timeout 900 python benchmarks/parity_report.py --md gpurun_out/r2_parity_report_a.md > gpurun_out/r2_parity_a.log 2>&1
cat gpurun_out/r2_tma_bo1.log gpurun_out/r2_tma_bo0.log gpurun_out/r2_tma_off.log gpurun_out/r2_t6.log; tail -12 gpurun_out/r2_parity_report_a.md
