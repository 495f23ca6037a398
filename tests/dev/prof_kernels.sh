# ncu --set full captures of the dominant kernels (one train step of config C); run under gpurun
N="ncu --set full --import-source on --clock-control none"
timeout 300 $N -k regex:tc_gate_bwd_kernel -s 10 -c 1 -f -o gpurun_out/r01_gate_bwd python tests/dev/prof_step.py 1 > gpurun_out/p3.log 2>&1
timeout 300 $N -k regex:tc_layer_kernel -s 12 -c 1 -f -o gpurun_out/r01_layer python tests/dev/prof_step.py 1 >> gpurun_out/p3.log 2>&1
timeout 300 $N -k regex:tc_dxw_kernel -s 10 -c 1 -f -o gpurun_out/r01_dxw python tests/dev/prof_step.py 1 >> gpurun_out/p3.log 2>&1
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tests/dev/prof_step.py 1 >> gpurun_out/p3.log 2>&1
ls -la gpurun_out/ | tail -6
