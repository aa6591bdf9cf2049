# final record run, part B: ncu launch list of the final code.  One iteration with the MAIN phases (Gmain + Dmain: 14 of every 16
# iterations are exactly this; --start-idx 1), because the all-phase list costs 14 GPU-minutes (profiles/r2_launches_by_kernel.csv
# has it for the record-run commit).
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_34
timeout 900 ncu --nvtx --nvtx-include "sgb_timed" --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_main.csv python bench.py --no-graphs --steps 1 --warmup 3 --lean --start-idx 1 > ${O}_ncu_launches.log 2>&1
find gpurun_out -type f -size +8M -print -delete
du -sh gpurun_out; wc -l ${O}_launches_main.csv
