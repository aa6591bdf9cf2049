# the record run of round 2: GPU tests, default bench (all arms), reference arm, op table vs the reference GPU path,
# parity report, ncu launch list + conv metrics + --set full of the hot kernels
set -x
mkdir -p gpurun_out
date +%s > gpurun_out/r2_20_t0
( time timeout 1500 python -m pytest tests -m gpu -q --tb=line -rf --durations=15 2>&1 | tail -45 ) > gpurun_out/r2_20_tests.log 2>&1
date +%s > gpurun_out/r2_20_t1
( time timeout 1500 python bench.py --breakdown gpurun_out/r2_20_bd.json ) > gpurun_out/r2_20_bench.log 2>&1
date +%s > gpurun_out/r2_20_t2
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_20_bench_ref.log 2>&1
timeout 1200 python benchmarks/vs_reference.py --out gpurun_out/r2_20_vs_reference.jsonl --md gpurun_out/r2_20_vs_reference.md > gpurun_out/r2_20_vs_reference.log 2>&1
timeout 900 python benchmarks/parity_report.py --md gpurun_out/r2_20_parity_report.md > gpurun_out/r2_20_parity.log 2>&1
date +%s > gpurun_out/r2_20_t3
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_20_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > gpurun_out/r2_20_ncu_launches.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" -k regex:"conv_halo_kernel|conv_tma|conv_wgrad|conv1x1_small|conv_umma|conv_simt" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_20_conv_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > gpurun_out/r2_20_ncu_conv.log 2>&1
date +%s > gpurun_out/r2_20_t4
CASES="fwd_f32_c64_256_n32,fwd_f32_c512_32,fwd_f16_c32_1024,fwd_f16_c64_512,wgrad_f32_c64_256,wgrad_f32_c512_32,wgrad_f16_c32_1024,convT_s2_f32_c128_128,conv_s2_f32_c64_256,fir_f16_c128_256,fir_f32_c64_256,up2_f16_c128_128,down2_f32_c64_256,bias_act_f16_c128_256"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > gpurun_out/r2_20_prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"halo|conv_tma|wgrad|upfirdn|bias_act" -c 40 -o gpurun_out/r2_20_prof python benchmarks/prof_shapes.py --reps 1 --cases $CASES > gpurun_out/r2_20_prof_ncu.log 2>&1
date +%s > gpurun_out/r2_20_t5
cat gpurun_out/r2_20_tests.log | tail -25; tail -c 3000 gpurun_out/r2_20_bench.log; tail -c 800 gpurun_out/r2_20_bench_ref.log; cat gpurun_out/r2_20_prof_plain.log
