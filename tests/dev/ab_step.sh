# in-call A/B of two library builds on one box: A = wavenet_b200/libwavenet_b200_A.so (baseline), B = the current build
for i in 1 2; do
  echo "A: $(WN_LIB_PATH=wavenet_b200/libwavenet_b200_A.so python tests/dev/time_step.py fp16x2 8 | tail -1)"
  echo "B: $(python tests/dev/time_step.py fp16x2 8 | tail -1)"
done
