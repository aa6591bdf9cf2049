# the record run of round 2 (second attempt: r2_run20 produced > 64 MiB of ncu reports and nothing came back):
# reports stay in /tmp on the box, only CSV exports travel
set -x
mkdir -p gpurun_out
date +%s > gpurun_out/r2_21_t0
( time timeout 1500 python -m pytest tests -m gpu -q --tb=line -rf 2>&1 | tail -15 ) > gpurun_out/r2_21_tests.log 2>&1
( time timeout 1500 python bench.py --breakdown gpurun_out/r2_21_bd.json ) > gpurun_out/r2_21_bench.log 2>&1
date +%s > gpurun_out/r2_21_t1
SGB_FUSED_CONV_MAIN=tma timeout 600 python bench.py --lean > gpurun_out/r2_21_bench_fusedmain.log 2>&1
SGB_FUSED_CONV_MAIN=tma timeout 600 python bench.py --lean --workload f1024 > gpurun_out/r2_21_bench_fusedmain_f1024.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_21_bench_ref.log 2>&1
date +%s > gpurun_out/r2_21_t2
timeout 1200 python benchmarks/vs_reference.py --out gpurun_out/r2_21_vs_reference.jsonl --md gpurun_out/r2_21_vs_reference.md > gpurun_out/r2_21_vs_reference.log 2>&1
date +%s > gpurun_out/r2_21_t3
timeout 900 python benchmarks/parity_report.py --md gpurun_out/r2_21_parity_report.md > gpurun_out/r2_21_parity.log 2>&1
date +%s > gpurun_out/r2_21_t4
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_21_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > gpurun_out/r2_21_ncu_launches.log 2>&1
date +%s > gpurun_out/r2_21_t5
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" -k regex:"conv_halo_kernel|conv_tma|conv_wgrad|conv1x1_small|conv_umma|conv_simt" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_21_conv_launches.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean > gpurun_out/r2_21_ncu_conv.log 2>&1
date +%s > gpurun_out/r2_21_t6
CASES="fwd_f32_c64_256_n32,fwd_f32_c512_32,fwd_f16_c32_1024,fwd_f16_c64_512,wgrad_f32_c64_256,wgrad_f32_c512_32,wgrad_f16_c32_1024,convT_s2_f32_c128_128,conv_s2_f32_c64_256,fir_f16_c128_256,fir_f32_c64_256,up2_f16_c128_128,down2_f32_c64_256,bias_act_f16_c128_256"
python benchmarks/prof_shapes.py --reps 3 --cases $CASES > gpurun_out/r2_21_prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"halo|conv_tma|wgrad|upfirdn|bias_act" -c 30 -o /tmp/r2_21_prof python benchmarks/prof_shapes.py --reps 1 --cases $CASES > gpurun_out/r2_21_prof_ncu.log 2>&1
ncu -i /tmp/r2_21_prof.ncu-rep --page raw --csv > gpurun_out/r2_21_prof_raw.csv 2>/dev/null
ls -la /tmp/r2_21_prof.ncu-rep
date +%s > gpurun_out/r2_21_t7
du -sh gpurun_out
tail -3 gpurun_out/r2_21_tests.log; for f in gpurun_out/r2_21_bench*.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); s=d.get('secondary') or {}
        print(sys.argv[1], d.get('config',{}).get('workload','')[:20], d['value'], d['ms_per_step'], 'f1024', s.get('value'), s.get('ms_per_step'))
PY
done
