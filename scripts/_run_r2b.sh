mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gpu_tests_r2b.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/gpu_tests_r2b.log
