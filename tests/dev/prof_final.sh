# final evidence pass of a round (under gpurun): GPU test suite, default bench line, launch list with DRAM bytes, ncu --set full
# of the per-layer kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final_launches_dram.csv python tests/dev/prof_step.py 2 fp16x2 > gpurun_out/p_final.log 2>&1
N="ncu --set full --import-source on --clock-control none"
timeout 300 $N -k regex:tcs_layer_kernel -s 12 -c 1 -f -o gpurun_out/final_tcs_layer python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_final.log 2>&1
timeout 300 $N -k regex:tcs_dxw_kernel -s 12 -c 1 -f -o gpurun_out/final_tcs_dxw python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_final.log 2>&1
timeout 300 $N -k regex:tcs_gate_bwd_kernel -s 12 -c 1 -f -o gpurun_out/final_tcs_gate_bwd python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_final.log 2>&1
timeout 300 $N -k regex:"tcs_gemm_kernel" -s 0 -c 1 -f -o gpurun_out/final_tcs_gemm_skip python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_final.log 2>&1
ls -la gpurun_out | grep final_
