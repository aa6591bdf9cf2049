set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q -x -k "cuda_graph or config_A" 2>&1 | tail -60 > gpurun_out/r2_t3_fail.log
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "upfirdn" 2>&1 | tail -30 > gpurun_out/r2_t3_fir.log
timeout 900 python benchmarks/parity_report.py --md gpurun_out/r2_parity_report.md > gpurun_out/r2_parity.log 2>&1
FIR="fir_f16_c128_256,fir_f16_c32_1024,fir_f32_c64_256,down2_f16_c32_1024,down2_f32_c64_256,up2_f16_c128_128,up2_f32_c64_128"
python benchmarks/prof_shapes.py --reps 5 --cases $FIR > gpurun_out/r2_fir_new.log 2>&1
SGB_FIR_OLD=1 python benchmarks/prof_shapes.py --reps 5 --cases $FIR > gpurun_out/r2_fir_old.log 2>&1
timeout 300 python benchmarks/ref_harness.py --backend reference --workload ffhq256 --fp32-mode strict --cudnn-benchmark --steps 4 > gpurun_out/r2_rh_ref_ffhq_strict.log 2>&1
CASES="fwd_f16_c32_1024,fwd_f16_c64_512,fwd_f32_c64_256_n32,wgrad_f16_c32_1024,wgrad_f16_c64_512,wgrad_f32_c64_256"
python benchmarks/prof_shapes.py --reps 2 --cases $CASES > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"halo|wgrad" -c 14 -o gpurun_out/r2_prof_conv python benchmarks/prof_shapes.py --reps 2 --cases $CASES > gpurun_out/r2_prof_ncu.log 2>&1
tail -20 gpurun_out/r2_t3_fail.log gpurun_out/r2_t3_fir.log gpurun_out/r2_fir_new.log gpurun_out/r2_fir_old.log gpurun_out/r2_prof_plain.log
