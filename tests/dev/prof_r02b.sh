# ncu --set full of the three per-layer fp16x2 kernels (one launch each, mid-stack layer) + step timing
N="ncu --set full --import-source on --clock-control none"
python tests/dev/time_step.py fp16x2 5 2>&1 | tail -1
python tests/dev/grad_err.py 2>&1 | grep -A1 "C 1x4200"
timeout 300 $N -k regex:tcs_layer_kernel -s 12 -c 1 -f -o gpurun_out/r02_tcs_layer python tests/dev/prof_step.py 1 fp16x2 > gpurun_out/p_r02b.log 2>&1
timeout 300 $N -k regex:tcs_dxw_kernel -s 12 -c 1 -f -o gpurun_out/r02_tcs_dxw python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_r02b.log 2>&1
timeout 300 $N -k regex:tcs_gate_bwd_kernel -s 12 -c 1 -f -o gpurun_out/r02_tcs_gate_bwd python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_r02b.log 2>&1
ls -la gpurun_out | grep r02_tcs
