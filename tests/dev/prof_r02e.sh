# sanitizer passes over the 4-CTA variant of gen_kernel_v6 (under gpurun)
S="compute-sanitizer --print-limit 3"
for tool in memcheck racecheck synccheck; do
  timeout 900 $S --tool $tool python tests/dev/sanitize_small.py gen6_4 > gpurun_out/r02e_san_${tool}_gen6_4.log 2>&1
  echo "$tool gen6_4 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02e_san_${tool}_gen6_4.log)"
done
python tests/dev/time_gen6.py 500 300 448 2>&1 | tail -4
