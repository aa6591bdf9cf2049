# dev run: role ablations of the halo kernel on the small 512-channel layers and the stride-2 forms (SGB_HALO_DEBUG bits:
# 1 no epilogue stores, 2 no patch loads, 4 no MMAs, 8 no weight loads, 64 no TF32 rounding pass)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_26
C="small_s1_16_n4,small_s2_8_n4,small_T2_8_n4,small_s1_16_n32,small_s2_16_n32,conv_s2_f32_c64_256,convT_s2_f32_c128_128,fwd_f32_c512_32"
for dbg in 0 64 2 66 4 8 1 79; do
  echo "== SGB_HALO_DEBUG=$dbg" >> ${O}_ablate.log
  SGB_HALO_DEBUG=$dbg python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $C >> ${O}_ablate.log 2>&1
done
echo "== SGB_HALO_PARBN=0" >> ${O}_ablate.log
SGB_HALO_PARBN=0 python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $C >> ${O}_ablate.log 2>&1
cat ${O}_ablate.log
