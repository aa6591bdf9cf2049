set -x
mkdir -p gpurun_out
timeout 900 python benchmarks/parity_report.py --md gpurun_out/r2_parity_report.md > gpurun_out/r2_parity.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t2_all.log
timeout 1200 python benchmarks/vs_reference.py --out gpurun_out/r2_vs_reference.jsonl --md gpurun_out/r2_vs_reference.md > gpurun_out/r2_vs_reference.log 2>&1
( time timeout 1200 python bench.py --breakdown gpurun_out/r2_bd_ffhq_1.json ) > gpurun_out/r2_bench_1.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_bench_ref_1.log 2>&1
tail -3 gpurun_out/r2_parity.log gpurun_out/r2_t2_all.log gpurun_out/r2_bench_1.log gpurun_out/r2_bench_ref_1.log
