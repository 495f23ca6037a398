# ncu --set full captures of the dominant kernels (one train step of config C); run under gpurun
N="ncu --set full --import-source on --clock-control none"
timeout 300 $N -k regex:tc_gate_bwd_kernel -s 10 -c 1 -f -o gpurun_out/r01_gate_bwd python tests/prof_step.py 1 > gpurun_out/p3.log 2>&1
timeout 300 $N -k regex:tc_layer_kernel -s 12 -c 1 -f -o gpurun_out/r01_layer python tests/prof_step.py 1 >> gpurun_out/p3.log 2>&1
timeout 300 $N -k regex:tc_gemm_kernel -s 10 -c 1 -f -o gpurun_out/r01_dx python tests/prof_step.py 1 >> gpurun_out/p3.log 2>&1
timeout 300 $N -k regex:tc_wgrad_kernel -s 10 -c 1 -f -o gpurun_out/r01_dw1 python tests/prof_step.py 1 >> gpurun_out/p3.log 2>&1
ls -la gpurun_out/ | tail -8
