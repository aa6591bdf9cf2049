set -x
mkdir -p gpurun_out
CASES="fwd_f16_c32_1024,fwd_f16_c64_512,fwd_f32_c64_256_n32"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > gpurun_out/r2_prof_tma_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_tma" -c 6 -o gpurun_out/r2_prof_tma python benchmarks/prof_shapes.py --reps 2 --cases $CASES > gpurun_out/r2_prof_tma_ncu.log 2>&1
cat gpurun_out/r2_prof_tma_plain.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -40 > gpurun_out/r2_t8.log; cat gpurun_out/r2_t8.log
