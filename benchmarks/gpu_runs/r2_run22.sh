# the record run of round 2 (third attempt: r2_run20 / r2_run21 left > 64 MiB in gpurun_out -- a 580 MB parity golden -- and nothing
# came back).  Reports stay in /tmp on the box; only logs / CSV exports travel, anything above 8 MB is dropped at the end.
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_22
date +%s > ${O}_t0
python benchmarks/measure_peaks.py ${O}_peaks.json > /dev/null 2>&1
( time timeout 1500 python bench.py --breakdown ${O}_bd.json ) > ${O}_bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > ${O}_bench_ref.log 2>&1
date +%s > ${O}_t1
timeout 1200 python benchmarks/vs_reference.py --out ${O}_vs_reference.jsonl --md ${O}_vs_reference.md > ${O}_vs_reference.log 2>&1
date +%s > ${O}_t2
timeout 900 python benchmarks/parity_report.py --md ${O}_parity_report.md > ${O}_parity.log 2>&1
rm -f /tmp/sgb200_parity_config_a_golden.npz
date +%s > ${O}_t3
timeout 600 python benchmarks/ref_harness.py --backend sgb200 --workload sg2attent256 --steps 8 > ${O}_rh_sgb_attent.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend reference --workload sg2attent256 --steps 8 --cudnn-benchmark > ${O}_rh_ref_attent.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend sgb200 --workload f1024 --steps 16 > ${O}_rh_sgb_f1024.log 2>&1
timeout 600 python benchmarks/ref_harness.py --backend reference --workload f1024 --steps 16 --cudnn-benchmark > ${O}_rh_ref_f1024.log 2>&1
date +%s > ${O}_t4
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > ${O}_ncu_launches.log 2>&1
date +%s > ${O}_t5
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" -k regex:"conv_halo_kernel|conv_tma|conv_wgrad|conv1x1_small|conv_umma|conv_simt" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file ${O}_conv_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > ${O}_ncu_conv.log 2>&1
date +%s > ${O}_t6
CASES="fwd_f32_c64_256_n32,fwd_f32_c512_32,fwd_f16_c32_1024,fwd_f16_c64_512,wgrad_f32_c64_256,wgrad_f32_c512_32,wgrad_f16_c32_1024,wgrad_f16_c64_512,convT_s2_f32_c128_128,conv_s2_f32_c64_256,fir_f16_c128_256,fir_f32_c64_256,fir_f16_c32_1024,up2_f16_c128_128,down2_f32_c64_256,bias_act_f16_c128_256"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > ${O}_prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"halo|conv_tma|wgrad|upfirdn|fir_sep|bias_act" -c 32 -o /tmp/r2_22_prof python benchmarks/prof_shapes.py --reps 1 --cases $CASES > ${O}_prof_ncu.log 2>&1
ncu -i /tmp/r2_22_prof.ncu-rep --page raw --csv > ${O}_prof_raw.csv 2>/dev/null
date +%s > ${O}_t7
timeout 600 python benchmarks/op_sweep.py --out ${O}_op_sweep.jsonl > ${O}_op_sweep.log 2>&1
date +%s > ${O}_t8
find gpurun_out -type f -size +8M -print -delete
du -sh gpurun_out
