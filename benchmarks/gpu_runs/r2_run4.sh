set -x
mkdir -p gpurun_out
timeout 900 python benchmarks/parity_report.py --md gpurun_out/r2_parity_report_rna.md > gpurun_out/r2_parity_rna.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_t4_all.log
timeout 600 python bench.py --lean > gpurun_out/r2_bench_rna.log 2>&1
SGB_HALO_DEBUG=64 timeout 600 python bench.py --lean > gpurun_out/r2_bench_trunc.log 2>&1
tail -5 gpurun_out/r2_t4_all.log; cat gpurun_out/r2_parity_report_rna.md; tail -c 600 gpurun_out/r2_bench_rna.log; tail -c 600 gpurun_out/r2_bench_trunc.log
