set -x
mkdir -p gpurun_out
SGB_TMA_FORCE=1 SGB_TMA_BO=1 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_tma_bo1.log 2>&1
SGB_TMA_FORCE=1 SGB_TMA_BO=0 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_tma_bo0.log 2>&1
cat gpurun_out/r2_tma_bo1.log gpurun_out/r2_tma_bo0.log
