mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_adam.py tests/test_gpu_losshead.py -m gpu -x -q > gpurun_out/gpu_tests_r2c.log 2>&1; echo "tests1 rc=$?"; tail -25 gpurun_out/gpu_tests_r2c.log
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_geometry.py tests/test_gpu_handnet.py -m gpu -x -q --durations=5 > gpurun_out/gpu_tests_r2c2.log 2>&1; echo "tests2 rc=$?"; tail -40 gpurun_out/gpu_tests_r2c2.log
for knob in "OBMAN_CONV64=0" "OBMAN_CONV64=1"; do
  echo "-- $knob"
  env $knob AB_B=256 AB_REPS=10 timeout 120 python scripts/ab_conv.py 2>&1 | grep -E "c64->64"
done
