mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_handnet.py tests/test_gpu_bn_train.py -q -x 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-secondary --no-gpu-eager --quick > gpurun_out/bench_r2ai.json 2> gpurun_out/bench_r2ai.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2ai.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'), 'traffic', r.get('traffic'), 'launches', d.get('gpu_launches'))
PY
OBMAN_TEST_NOOP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-gpu-eager --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('repeat', round(d['ms_per_step'],3))"
