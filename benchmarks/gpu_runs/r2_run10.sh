set -x
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown gpurun_out/r2_bd_u.json > gpurun_out/r2_bench_u2.log 2>&1
SGB_TMA=0 timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --breakdown gpurun_out/r2_bd_u_notma.json > gpurun_out/r2_bench_u2_notma.log 2>&1
