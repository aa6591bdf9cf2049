# dev run: where the time of the small 512-channel layers goes (config-f at 4 images per GPU), in-bench A/B of the kys weight gradient
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_24
S4="small_s1_8_n4,small_s1_16_n4,small_s1_32_n4,small_s2_8_n4,small_s2_16_n4,small_T1_16_n4,small_T2_8_n4"
S32="small_s1_8_n32,small_s1_16_n32,small_s1_32_n32,small_s2_8_n32,small_s2_16_n32,small_T1_16_n32,small_T2_8_n32"
python benchmarks/prof_shapes.py --reps 5 --graph --inner 10 --cases $S4,$S32 > ${O}_small_graph.log 2>&1
python benchmarks/prof_shapes.py --reps 5 --cases $S4 > ${O}_small_eager.log 2>&1
cat ${O}_small_graph.log ${O}_small_eager.log
P="small_s1_16_n4,small_s2_8_n4,small_T2_8_n4,small_s1_16_n32"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|pack_weights" -c 8 -o /tmp/r2_24_small python benchmarks/prof_shapes.py --reps 1 --cases $P > ${O}_small_ncu.log 2>&1
ncu -i /tmp/r2_24_small.ncu-rep --page raw --csv > ${O}_small_raw.csv 2>/dev/null
for i in 0 1 2 3 4 5 6 7; do python benchmarks/ncu_hot_lines.py /tmp/r2_24_small.ncu-rep --launch $i --top 22 > ${O}_small_hot_$i.txt 2>&1; done
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-strict --no-callers --no-secondary --breakdown ${O}_bd.json > ${O}_bench.log 2>&1
SGB_WGRAD_KYS=0 timeout 600 python bench.py --lean > ${O}_bench_kys0.log 2>&1
for f in ${O}_bench.log ${O}_bench_kys0.log; do python - $f <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d=json.loads(ln); print(sys.argv[1], d['value'], d['ms_per_step'])
PY
done
find gpurun_out -type f -size +8M -print -delete
du -sh gpurun_out
