set -x
mkdir -p gpurun_out
CASES="fir_f32_c64_256,fir_f16_c32_1024,bias_act_f16_c128_256"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > gpurun_out/r2_prof_fir_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"upfirdn|bias_act" -c 6 -o gpurun_out/r2_prof_fir python benchmarks/prof_shapes.py --reps 2 --cases $CASES > gpurun_out/r2_prof_fir_ncu.log 2>&1
cat gpurun_out/r2_prof_fir_plain.log
