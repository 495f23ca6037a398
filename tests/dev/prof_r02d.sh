# round-2 evidence for gen_kernel_v6 (under gpurun): ncu --set full of one launch (1920 streams, 200 steps), sanitizer passes
N="ncu --set full --import-source on --clock-control none"
timeout 600 $N -k regex:gen_kernel_v6 -c 1 -f -o gpurun_out/r02d_gen_v6 python tests/dev/prof_gen256.py 1920 200 > gpurun_out/p_r02d.log 2>&1
tail -2 gpurun_out/p_r02d.log
S="compute-sanitizer --print-limit 3"
for tool in memcheck racecheck synccheck; do
  timeout 900 $S --tool $tool python tests/dev/sanitize_small.py gen6 > gpurun_out/r02d_san_${tool}_gen6.log 2>&1
  echo "$tool gen6 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02d_san_${tool}_gen6.log)"
done
