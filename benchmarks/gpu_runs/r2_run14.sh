set -x
mkdir -p gpurun_out
FIR="fir_f16_c128_256,fir_f16_c32_1024,fir_f32_c64_256,down2_f16_c32_1024,down2_f32_c64_256,up2_f16_c128_128,up2_f32_c64_128"
SGB_FIR_V2=1 python benchmarks/prof_shapes.py --reps 5 --cases $FIR > gpurun_out/r2_fir_v2.log 2>&1
SGB_FIR_V2=0 python benchmarks/prof_shapes.py --reps 5 --cases $FIR > gpurun_out/r2_fir_v1.log 2>&1
echo V2; cat gpurun_out/r2_fir_v2.log; echo V1; cat gpurun_out/r2_fir_v1.log
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_networks.py -m gpu -q --tb=short -rf -x 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20
