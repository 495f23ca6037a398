# round-2 compute-sanitizer evidence (under gpurun): memcheck / racecheck / synccheck over one fp16x2 + one tf32 train step of
# the fused shape and a few steps of the cluster + single-CTA generators (tests/dev/sanitize_small.py)
S="compute-sanitizer --print-limit 3"
for tool in memcheck racecheck; do
  for what in train gen; do
    timeout 600 $S --tool $tool python tests/dev/sanitize_small.py $what > gpurun_out/r02_san_${tool}_${what}.log 2>&1
    echo "$tool $what rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_san_${tool}_${what}.log)"
  done
done
timeout 600 $S --tool synccheck python tests/dev/sanitize_small.py gen > gpurun_out/r02_san_synccheck_gen.log 2>&1
echo "synccheck gen rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r02_san_synccheck_gen.log)"
# synccheck of the training kernels: everything except tcs_gate_bwd_kernel, then that kernel alone
timeout 600 $S --tool synccheck --kernel-name-exclude kns=tcs_gate_bwd python tests/dev/sanitize_small.py train > gpurun_out/r02_san_synccheck_train_other.log 2>&1
echo "synccheck train (all but tcs_gate_bwd) rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r02_san_synccheck_train_other.log)"
timeout 600 $S --tool synccheck --kernel-name kns=tcs_gate_bwd python tests/dev/sanitize_small.py train > gpurun_out/r02_san_synccheck_train_gate_bwd.log 2>&1
echo "synccheck train (tcs_gate_bwd only) rc=$? $(grep 'ERROR SUMMARY' gpurun_out/r02_san_synccheck_train_gate_bwd.log)"
grep -m3 -A6 "Barrier error" gpurun_out/r02_san_synccheck_train_gate_bwd.log | head -30
