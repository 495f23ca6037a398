# round-2 evidence run (under gpurun): full GPU test suite, default bench line, launch list with DRAM bytes of the fp16x2
# train step, ncu --set full of the N=256 GEMMs and of the generator kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
N="ncu --set full --import-source on --clock-control none"
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_dram_fp16x2.csv python tests/dev/prof_step.py 2 fp16x2 > gpurun_out/p_r02.log 2>&1
timeout 300 $N -k regex:"tcs_gemm_kernel" -s 0 -c 3 -f -o gpurun_out/r02_tcs_gemm python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_r02.log 2>&1
timeout 300 $N -k regex:"tcs_wgrad_kernel" -s 0 -c 3 -f -o gpurun_out/r02_tcs_wgrad python tests/dev/prof_step.py 1 fp16x2 >> gpurun_out/p_r02.log 2>&1
timeout 300 $N -k regex:gen_kernel_v4 -c 1 -f -o gpurun_out/r02_gen_v4 python tests/dev/prof_gen.py 1 300 >> gpurun_out/p_r02.log 2>&1
timeout 300 $N -k regex:gen_kernel_v3 -c 1 -f -o gpurun_out/r02_gen_v3_256 python tests/dev/prof_gen.py 256 300 >> gpurun_out/p_r02.log 2>&1
ls -la gpurun_out/ | grep r02
