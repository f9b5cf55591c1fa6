mkdir -p gpurun_out
OBMAN_RAYCAST_PACKED=1 timeout 200 python scripts/time_raycast.py 2>&1 | grep "^raycast"
OBMAN_RAYCAST_PACKED=0 timeout 200 python scripts/time_raycast.py 2>&1 | grep "^raycast"
timeout 600 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_losshead.py tests/test_gpu_handnet.py -q -x 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-gpu-eager --quick > gpurun_out/bench_r2ac.json 2> gpurun_out/bench_r2ac.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2ac.json').read().strip().splitlines()[-1])
r = d.get('roofline', {})
print(round(d['ms_per_step'], 3), round(d['value'], 1), 'pipe', r.get('tensor_pipe_frac'), 'gemm_ms', r.get('gemm_ms_per_step'))
PY
