# final record run of round 2, part A (final code): smoke, GPU tests, all bench arms, op table vs the reference GPU path, parity report,
# unchanged callers on configs C / D, conv launch list, --set full of the hot kernels.  Part B (r2_run34.sh) = the full launch list.
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_33
date +%s > ${O}_t0
python __graft_entry__.py smoke > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q --tb=line -rf 2>&1 | tail -8 ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
date +%s > ${O}_t1
( time timeout 1500 python bench.py --breakdown ${O}_bd.json ) > ${O}_bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > ${O}_bench_ref.log 2>&1
date +%s > ${O}_t2
timeout 1200 python benchmarks/vs_reference.py --out ${O}_vs_reference.jsonl --md ${O}_vs_reference.md > ${O}_vs_reference.log 2>&1
timeout 900 python benchmarks/parity_report.py --md ${O}_parity_report.md > ${O}_parity.log 2>&1
rm -f /tmp/sgb200_parity_config_a_golden.npz
date +%s > ${O}_t3
timeout 600 python benchmarks/ref_harness.py --backend sgb200 --workload sg2attent256 --steps 8 > ${O}_rh_sgb_attent.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend sgb200 --workload f1024 --steps 16 > ${O}_rh_sgb_f1024.log 2>&1
date +%s > ${O}_t4
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" -k regex:"conv_halo_kernel|conv_tma|conv_wgrad|conv1x1_small|conv_umma|conv_simt" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file ${O}_conv_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > ${O}_ncu_conv.log 2>&1
date +%s > ${O}_t5
CASES="fwd_f32_c64_256_n32,fwd_f32_c512_32,fwd_f16_c32_1024,fwd_f16_c64_512,small_s1_16_n4,wgrad_f32_c64_256,wgrad_f32_c512_32,wgrad_f16_c32_1024,wgrad_f16_c64_512,convT_s2_f32_c128_128,conv_s2_f32_c64_256,fir_f16_c128_256,fir_f32_c64_256,fir_f16_c32_1024,up2_f16_c128_128,down2_f32_c64_256,bias_act_f16_c128_256"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > ${O}_prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"halo|conv_tma|wgrad|upfirdn|fir_sep|bias_act" -c 34 -o /tmp/r2_33_prof python benchmarks/prof_shapes.py --reps 1 --cases $CASES > ${O}_prof_ncu.log 2>&1
ncu -i /tmp/r2_33_prof.ncu-rep --page raw --csv > ${O}_prof_raw.csv 2>/dev/null
date +%s > ${O}_t6
find gpurun_out -type f -size +8M -print -delete
du -sh gpurun_out
tail -c 1500 ${O}_bench.log
