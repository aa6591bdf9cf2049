set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_ops_gpu.py 2>&1 | tail -40 > gpurun_out/r2_t1_new.log
timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_conv_umma_gpu.py tests/test_fused_conv_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_t1_ops.log
timeout 600 python bench.py --breakdown gpurun_out/r2_bd_ffhq_0.json > gpurun_out/r2_bench_ffhq_0.log 2>&1
timeout 600 python bench.py --workload f1024 --no-cpu-baseline --breakdown gpurun_out/r2_bd_f1024_0.json > gpurun_out/r2_bench_f1024_0.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend sgb200 --workload ffhq256 > gpurun_out/r2_rh_sgb_ffhq.log 2>&1
timeout 900 python benchmarks/ref_harness.py --backend reference --workload ffhq256 > gpurun_out/r2_rh_ref_ffhq.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend reference --workload ffhq256 --cudnn-benchmark > gpurun_out/r2_rh_ref_ffhq_cb.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend sgb200 --workload f1024 > gpurun_out/r2_rh_sgb_f1024.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend reference --workload f1024 --cudnn-benchmark > gpurun_out/r2_rh_ref_f1024_cb.log 2>&1
tail -3 gpurun_out/r2_*.log
