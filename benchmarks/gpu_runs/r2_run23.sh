# dev run: weight gradient with the filter rows stacked in M (tf32: conv_wgrad_tf32_kernel kys mode; 16-bit: conv_wgrad_kys_kernel)
set -x
mkdir -p gpurun_out
O=gpurun_out/r2_23
timeout 600 python -m pytest tests/test_conv_umma_gpu.py -m gpu -q --tb=short -rf -x -k "wgrad or double_backward or gradients" 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30 > ${O}_t_wgrad.log; cat ${O}_t_wgrad.log
W="wgrad_f32_c64_256,wgrad_f16_c32_1024,wgrad_f16_c64_512,wgrad_f32_c512_32"
SGB_WGRAD_KYS=1 python benchmarks/prof_shapes.py --reps 5 --cases $W > ${O}_w_kys1.log 2>&1
SGB_WGRAD_KYS=0 python benchmarks/prof_shapes.py --reps 5 --cases $W > ${O}_w_kys0.log 2>&1
SGB_WGRAD_KYS=1 SGB_WGRAD_TH=8 python benchmarks/prof_shapes.py --reps 5 --cases $W > ${O}_w_kys1_th8.log 2>&1
echo KYS1; cat ${O}_w_kys1.log; echo KYS0; cat ${O}_w_kys0.log; echo KYS1_TH8; cat ${O}_w_kys1_th8.log
timeout 900 python -m pytest tests/test_networks.py tests/test_ref_callers_gpu.py -m gpu -q --tb=line -rf 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20 > ${O}_t_net.log; cat ${O}_t_net.log
timeout 600 python bench.py --lean > ${O}_bench.log 2>&1; tail -c 500 ${O}_bench.log
timeout 600 python bench.py --lean --workload f1024 > ${O}_bench_f1024.log 2>&1; tail -c 500 ${O}_bench_f1024.log
find gpurun_out -type f -size +8M -print -delete
