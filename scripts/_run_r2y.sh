mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q -x -k "gemm or decoder or linear" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_handnet.py tests/test_gpu_bn_train.py -q -x 2>&1 | tail -3
for P in 1 0; do
OBMAN_GEMM_PERSIST=$P timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-gpu-eager --quick --dump-launches gpurun_out/tc_launches_r2y_$P.txt > gpurun_out/bench_r2y_$P.json 2> gpurun_out/bench_r2y_$P.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_r2y_$P.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print('PERSIST=$P', round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'))
PY
grep "gemm M655872" gpurun_out/tc_launches_r2y_$P.txt | head
done
