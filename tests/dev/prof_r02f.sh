# final ncu captures of the tensor-core generator (under gpurun): 8-CTA clusters at 1920 streams, 4-CTA clusters at 4224
N="ncu --set full --import-source on --clock-control none"
timeout 600 $N -k regex:gen_kernel_v6 -c 1 -f -o gpurun_out/r02f_gen_v6_8cta python tests/dev/prof_gen256.py 1920 200 > gpurun_out/p_r02f.log 2>&1
tail -1 gpurun_out/p_r02f.log
timeout 600 $N -k regex:gen_kernel_v6 -c 1 -f -o gpurun_out/r02f_gen_v6_4cta python tests/dev/prof_gen256.py 4224 200 >> gpurun_out/p_r02f.log 2>&1
tail -1 gpurun_out/p_r02f.log
