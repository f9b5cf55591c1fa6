set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -x -q -m gpu > gpurun_out/cl_dense.log 2>&1; echo "dense rc=$?" 
tail -5 gpurun_out/cl_dense.log
for cl in 1 2 4; do
  OBMAN_GEMM_CLUSTER=$cl PROF_REPS=20 timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_cl$cl.txt 2>&1; echo "prof cl=$cl rc=$?"
  OBMAN_GEMM_CLUSTER=$cl timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cl$cl.json 2> gpurun_out/bench_cl$cl.err; echo "bench cl=$cl rc=$?"
  tail -c 1500 gpurun_out/bench_cl$cl.json
done
OBMAN_GEMM_CLUSTER=4 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_encoder.py -x -q -m gpu > gpurun_out/cl4_tests.log 2>&1; echo "cl4 tests rc=$?"
tail -3 gpurun_out/cl4_tests.log
timeout 400 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_handnet.py -x -q -m gpu > gpurun_out/cl2_tests.log 2>&1; echo "cl2 tests rc=$?"
tail -3 gpurun_out/cl2_tests.log
