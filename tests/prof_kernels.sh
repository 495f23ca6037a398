# ncu --set full captures of the per-layer backward kernels (one train step of config C); run under gpurun
N="ncu --set full --import-source on --clock-control none"
timeout 300 $N -k regex:tc_gemm_kernel -s 10 -c 2 -f -o gpurun_out/p2_gemm64 python tests/prof_step.py 1 > gpurun_out/p2.log 2>&1
timeout 300 $N -k regex:tc_wgrad_kernel -s 20 -c 3 -f -o gpurun_out/p2_wgrad python tests/prof_step.py 1 >> gpurun_out/p2.log 2>&1
ls -la gpurun_out/
