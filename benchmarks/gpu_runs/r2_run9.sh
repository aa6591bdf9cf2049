set -x
mkdir -p gpurun_out
SGB_TMA_FORCE=1 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_tma_u.log 2>&1
SGB_TMA=0 timeout 600 python benchmarks/experiments/tma_check.py > gpurun_out/r2_halo_u.log 2>&1
cat gpurun_out/r2_tma_u.log gpurun_out/r2_halo_u.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20 > gpurun_out/r2_t9.log; cat gpurun_out/r2_t9.log
timeout 600 python bench.py --lean > gpurun_out/r2_bench_u.log 2>&1; tail -c 400 gpurun_out/r2_bench_u.log
timeout 600 python bench.py --lean --workload f1024 > gpurun_out/r2_bench_u_f1024.log 2>&1; tail -c 400 gpurun_out/r2_bench_u_f1024.log
